// conv3x3_trunk.cu -- the 23 x 3 residual dense blocks of one L2-resident tile group as ONE persistent kernel,
// organised so that consecutive layer passes OVERLAP instead of draining the SM between them.
//
// What the earlier persistent kernel (conv3x3_body.cu) taught (gpurun_out/decompose.log, r1 launch list):
//   * swept over a whole 1080p frame, every layer pass re-reads its 64..192 input channels from HBM -- the
//     frame's dense-block buffer (213 MB) cannot live in the 126 MB L2, ncu shows 86 GB of DRAM reads per
//     frame, and the trunk runs at the HBM roofline, not the tensor roofline;
//   * shrinking the working set (fewer tiles per batch) made it SLOWER: every pass costs ~10 us of fixed
//     latency (drain the epilogue, grid-wide arrival counter, reload up to 110 KB of weights, refill the
//     ring) and small batches have short passes.
// This kernel removes the fixed cost, so that tile groups small enough for L2 pay off:
//
//   TMEM-RESIDENT BANDS.  A CTA owns at most 16 output rows (one or two bands of a 128-pixel strip); the
//   fp32 accumulators of ALL of them (16 row slots x 32 channels = 512 TMEM columns) stay in TMEM for the
//   whole pass.  No ring, no wrap: output row j of a band is slot slot0+j in every pass.
//   CHUNK-MAJOR SWEEPS.  Because the band is TMEM resident, the contraction can be ordered chunk by chunk:
//   for each 64-channel input chunk, stream the band's input rows once and accumulate.  The weights then
//   stream too -- one [3 dx][96][64] box set (36 KB) per chunk through a 3-deep ring -- so a pass never
//   waits for "its" weights and no pass needs more than 36 KB of them resident.
//   DEPENDENCIES ARRIVE LAST.  Within a dense block conv_k+1 differs from conv_k only by the 32 newest
//   input channels, and those sit in the LAST chunk.  While pass k's final sweep is still being drained,
//   stored and published, pass k+1 already sweeps chunk 0 (and 1): the producer only waits before loading
//   the chunk that holds another pass's output (ConvParams::need).  Only conv1 of the next block (a
//   single chunk that IS the previous block's output) waits exposed.
//   NEIGHBOUR PROGRESS WORDS, NO GRID BARRIER.  A band's input halo is written by the few CTAs that own
//   the adjacent bands.  Every CTA publishes "passes completed" in its own 128-byte line; the producer
//   warp polls the words of its halo neighbours (host-built list, one lane each) and of its own CTA.  A
//   slow CTA delays its neighbours, not all 148 SMs (the grid-wide counter cost 5-22k cycles of skew per pass).
//   The epilogue warps publish a pass with their own named barrier; TMA and MMA warps never stop.
//   TWO BAND SETS PER CTA (optional, NESR_B200_SETS=2; off by default).  The schedule can give each CTA two band sets far
//   apart in the strip sequence and the kernel then alternates (pass k, set 0), (pass k, set 1), (pass k+1, set 0) ... with
//   progress words and halo lists per (CTA, set), so that one set's publish -> acquire latency hides behind the other set's
//   work.  Measured SLOWER (6.7 vs 6.0 ms per group): the cross-CTA machinery costs only ~7 % in total, the three roles of a
//   CTA are the bound, and the second band costs two more halo rows (profiles/r1_trunk_experiments.txt).
//
// Roles (fold_roles.cuh explains the row fold itself): warp 0 TMA producer (weights + row slabs), warp 1
// MMA issuer, warps 2..9 epilogue (two groups alternating rows).  Launched cooperatively (grid <= #SMs).
#include <stdio.h>

#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace nesr {

namespace {

constexpr int COUT = 32;
constexpr int kThreads = 320;
constexpr int kSlabPx = 136;
constexpr int kSlabBytes = kSlabPx * 128;              // 17408
constexpr int kStages = 6;                             // activation slab ring
constexpr int kWStages = 3;                            // weight chunk ring
constexpr int kWBoxBytes = 3 * COUT * 128;             // one dx: [96 rows][64 ch] = 12288
constexpr int kWChunkBytes = 3 * kWBoxBytes;           // three dx boxes = 36864
constexpr int kSlots = 16;                             // TMEM row slots (512 / 32)
constexpr int kMaxBands = kTrunkMaxBands;
#if NESR_PROF
constexpr int kTracePasses = 48;
#define TS(k, pass) do { if ((pass) < kTracePasses) sh.ts[k][pass] = clock64(); } while (0)
#define EPI_T(var) const long long var = clock64()
#define EPI_ACC(k, pass, dt) do { if (threadIdx.x == 64 && (pass) < kTracePasses) sh.epi_acc[k][pass] += (dt); } while (0)
#else
#define TS(k, pass) do {} while (0)
#define EPI_T(var) do {} while (0)
#define EPI_ACC(k, pass, dt) do {} while (0)
#endif

constexpr int kMaxOps = 20;                            // TMA operations per slab row of a packed strip
struct BandInfo {                                      // one band of this CTA, derived once at kernel start
  int32_t rows, slot0, nop, full_strip;
  uint32_t row_bytes;
  int32_t op_px[kMaxOps];                              // flat pixel of the box at input row r0-1
  int32_t op_pitch[kMaxOps];                           // row pitch of its tile
  int32_t op_off[kMaxOps];                             // byte offset of the box inside the slab
  int32_t op_box[kMaxOps];                             // box size index: 8 << op_box pixels
};

struct Shared {
  uint64_t wfull[kWStages], wempty[kWStages];
  uint64_t full[kStages], empty[kStages];
  uint64_t tfull[kSlots], tempty[kSlots];
  uint32_t tmem_slot;
  int32_t nband, nband0;                                // bands of this CTA; the first nband0 are set 0, the rest set 1
  BandInfo band[kMaxBands];
  int32_t lane_px[kMaxBands][128];                     // flat pixel of (r0, x) of each MMA lane, or -1 (masked lane)
  int32_t lane_pitch[kMaxBands][128];
  int32_t lane_rows[kMaxBands][128];                   // band rows [0, lane_rows) belong to the lane's piece
#if NESR_PROF
  long long ts[6][kTracePasses];                       // per-pass time stamps of the dependency chain (debug_flags & 1024)
  long long epi_acc[4][kTracePasses];                  // epilogue warp 2: cycles in wait_tfull / tmem ld+zero+arrive / math+stores / rows
#endif
};

constexpr int kRingBytes = kWStages * kWChunkBytes + kStages * kSlabBytes;      // 110592 + 104448
constexpr int kSmemBytes = kRingBytes + static_cast<int>(sizeof(Shared)) + 1024;
static_assert(kSmemBytes <= 232448, "trunk kernel: shared memory budget");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* ptr, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
constexpr int kProgStride = 32;                        // one 128-byte line per CTA
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// The fields of a pass the TMA producer / MMA issuer need, fetched one pass ahead.
struct PassHead {
  int32_t cin, w_row0, src_sel, need0, need1, need2, dbg;
};
__device__ __forceinline__ PassHead load_head(const ConvParams* passes, int pass, int npass) {
  PassHead h{};
  if (pass < npass) {
    const ConvParams* p = passes + pass;
    h.cin = __ldg(&p->cin); h.w_row0 = __ldg(&p->w_row0); h.src_sel = __ldg(&p->src_sel);
    h.need0 = __ldg(&p->need[0]); h.need1 = __ldg(&p->need[1]); h.need2 = __ldg(&p->need[2]);
    h.dbg = __ldg(&p->debug_flags);
  }
  return h;
}

// One chunk sweep over one band, executed by the single MMA-issuing thread.  KS k-steps per (row, dx); FIRST: first
// sweep of the pass (wait until the epilogue has drained + zeroed a slot before its first MMA); LAST: last sweep (commit
// each output row's completion).  The thread is the bottleneck of a sweep (~120 instructions per row at one instruction
// per ~4-5 cycles against 684 cycles of MMA work), so everything that can be is a template parameter and interior rows
// (all three output rows inside the band) take a path without clamps.
template <int KS, bool FIRST, bool LAST>
__device__ __forceinline__ void sweep_band(Shared& sh, const int rows, const int slot0, const uint32_t tmem_base, const uint32_t hw,
                                           const uint32_t hi, const uint32_t a_lo0, const uint32_t w_lo, const uint32_t tparity,
                                           int& stage, uint32_t& phase, const bool mma_on) {
  constexpr uint32_t kSlabLo = kSlabBytes >> 4, kWBoxLo = kWBoxBytes >> 4;
  constexpr uint32_t kBlkLo = (COUT * 128) >> 4;                // one dy block of 32 weight rows
  const uint32_t id96 = umma_idesc_f16(hw, 3 * COUT);
  // input row -1: its slab, and (first sweep of the pass) output slot 0 drained + zeroed by the epilogue
  if (FIRST) mbar_wait(&sh.tempty[slot0], tparity ^ 1);
  mbar_wait(&sh.full[stage], phase);
  tc_fence_after();
  for (int i = -1; i <= rows; ++i) {
    uint32_t d, id, b_lo;
    if (i >= 1 && i + 1 <= rows - 1) {                          // interior: output rows i-1, i, i+1
      d = tmem_base + static_cast<uint32_t>(slot0 + i - 1) * COUT; id = id96; b_lo = w_lo;
    } else {
      const int lo = i - 1 < 0 ? 0 : i - 1;
      const int hi_row = i + 1 > rows - 1 ? rows - 1 : i + 1;
      d = tmem_base + static_cast<uint32_t>(slot0 + lo) * COUT;
      id = umma_idesc_f16(hw, static_cast<uint32_t>(COUT * (hi_row - lo + 1)));
      b_lo = w_lo + static_cast<uint32_t>(lo - (i - 1)) * kBlkLo;
    }
    const uint32_t a_lo = a_lo0 + stage * kSlabLo;
    if (mma_on) umma_f16_ksteps<KS>(d, a_lo, b_lo, hi, id);
    // while those run: is the next input row ready?  (slot i+2 is first touched by input row i+1)
    const int nstage = stage + 1 == kStages ? 0 : stage + 1;
    if (i < rows) {
      if (FIRST && i + 2 <= rows - 1) mbar_wait(&sh.tempty[slot0 + i + 2], tparity ^ 1);
      mbar_wait(&sh.full[nstage], nstage == 0 ? phase ^ 1 : phase);
      tc_fence_after();
    }
    if (mma_on) {
      umma_f16_ksteps<KS>(d, a_lo + 8, b_lo + kWBoxLo, hi, id);
      umma_f16_ksteps<KS>(d, a_lo + 16, b_lo + 2 * kWBoxLo, hi, id);
    }
    umma_commit(&sh.empty[stage]);                              // slab may be overwritten once these MMAs have read it
    if (LAST && i >= 1) umma_commit(&sh.tfull[slot0 + i - 1]);  // output row i-1 has all its contributions
    stage = nstage;
    if (stage == 0) phase ^= 1;
  }
}

// WMC (weight multicast): launched as clusters of two CTAs; each loads HALF of every weight chunk (48 of the 96 rows of each dx
// box) and multicasts it into both CTAs' weight rings, halving the weight bytes read from L2 (weights are ~18 % of the TMA
// bytes).  A ring slot is refilled once BOTH CTAs' MMAs have released it (multicast commit, barrier count 2).
template <bool WMC>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_trunk_kernel(const __grid_constant__ TrunkMaps maps, const ConvParams* __restrict__ passes, const int npass,
                     unsigned* __restrict__ prog) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wring = smem;
  uint8_t* ring = smem + kWStages * kWChunkBytes;
  Shared& sh = *reinterpret_cast<Shared*>(smem + kRingBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup: barriers, TMEM, band geometry (identical for every pass of the trunk) ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.full[0]); tma_prefetch_desc(&maps.full[1]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < kWStages; ++i) { mbar_init(&sh.wfull[i], 1); mbar_init(&sh.wempty[i], WMC ? 2 : 1); }
    for (int i = 0; i < kStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < kSlots; ++i) { mbar_init(&sh.tfull[i], 1); mbar_init(&sh.tempty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh.tmem_slot, 512);
    tmem_relinquish();
  }
  {
    const ConvParams& p0 = passes[0];
    const int band_begin = p0.cta_band_off[blockIdx.x];
    const int band_end = p0.cta_band_off[blockIdx.x + 1];
    const int nband = min(band_end - band_begin, kMaxBands);
    if (threadIdx.x == 0) {
      sh.nband = nband;
      sh.nband0 = min(__ldg(p0.trunk_split + blockIdx.x), nband);
      int slot0 = 0;
      for (int b = 0; b < nband; ++b) {
        const FoldBand band = p0.bands[band_begin + b];
        BandInfo& bi = sh.band[b];
        bi.rows = band.rows; bi.slot0 = slot0; bi.nop = 0; bi.full_strip = 0;
        slot0 += band.rows;
        uint32_t row_bytes = 0;
        for (int sgi = 0; sgi < band.nseg; ++sgi) {
          const FoldSeg sg = p0.segs[band.seg0 + sgi];
          const LevelGeom g = p0.tiles[sg.tile].lv[0];
          const int px = g.base + (band.r0 - 1 + sg.y0) * g.pitch + sg.x0 - 1;
          if (sg.width == kBlockPixels) {                        // a 128-pixel segment is always alone: one 136-pixel box
            bi.full_strip = 1;
            bi.op_px[0] = px; bi.op_pitch[0] = g.pitch; bi.op_off[0] = 0; bi.op_box[0] = 0; bi.nop = 1;
            break;
          }
          int n8 = (sg.width + 2 + 7) >> 3, done8 = 0;           // 8-pixel units incl. halo, largest boxes first
          for (int k = 3; k >= 0; --k)
            while (n8 - done8 >= (1 << k) && bi.nop < kMaxOps) {
              bi.op_px[bi.nop] = px + done8 * 8; bi.op_pitch[bi.nop] = g.pitch;
              bi.op_off[bi.nop] = sg.lane0 * 128 + done8 * 1024; bi.op_box[bi.nop] = k;
              ++bi.nop;
              done8 += 1 << k;
            }
          row_bytes += n8 * 1024;
        }
        bi.row_bytes = bi.full_strip ? kSlabBytes : row_bytes;
      }
    }
    if (threadIdx.x < 128) {
      const int m = threadIdx.x;
      for (int b = 0; b < nband; ++b) {
        const FoldBand band = p0.bands[band_begin + b];
        int px = -1, pitch = 0, nrow = 0;
        for (int sgi = 0; sgi < band.nseg; ++sgi) {
          const FoldSeg sg = p0.segs[band.seg0 + sgi];
          if (m >= sg.lane0 && m < sg.lane0 + sg.width) {
            const LevelGeom g = p0.tiles[sg.tile].lv[0];
            const int x = sg.x0 + (m - sg.lane0);
            if (x < g.w) { px = g.base + (band.r0 + sg.y0) * g.pitch + x; pitch = g.pitch; nrow = min(max(sg.h - band.r0, 0), band.rows); }
          }
        }
        sh.lane_px[b][m] = px;
        sh.lane_pitch[b][m] = pitch;
        sh.lane_rows[b][m] = nrow;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (WMC) cluster_sync_all();                                  // the peer's barriers exist before anything is multicast into them
  const uint32_t tmem_base = sh.tmem_slot;
#if NESR_PROF
  for (int i = threadIdx.x; i < 4 * kTracePasses; i += kThreads) (&sh.epi_acc[0][0])[i] = 0;
#endif
  if (warp >= 2 && warp < 6) {                                  // every MMA accumulates: start from zero
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = 0; c < 512; c += 16) tmem_st16_zero(t0 + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int nband = sh.nband, nband0 = sh.nband0;
  const int nsets = nband0 < nband ? 2 : 1;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    int stage = 0; uint32_t phase = 0;                          // slab ring
    int ws = 0; uint32_t wphase = 0;                            // weight ring
    unsigned known[2] = {0, 0};                                 // passes known to be complete on every halo neighbour, per set
    const int plane_px = __ldg(&passes[0].src_plane_px);
    const uint64_t keep = l2_policy_evict_last();               // dense-block activations and weights: stay in L2
    const unsigned* my_dep[2];
    for (int st = 0; st < 2; ++st)
      my_dep[st] = prog + static_cast<size_t>(__ldg(passes[0].trunk_deps + (blockIdx.x * 2 + st) * kTrunkMaxDeps + lane)) * kProgStride;
    PassHead h = load_head(passes, 0, npass);
    for (int pass = 0; pass < npass; ++pass) {
      const PassHead nh = load_head(passes, pass + 1, npass);   // next pass, fetched early
      const int nchunk = (h.cin + kChunkChannels - 1) / kChunkChannels;
      const CUtensorMap* amap = &maps.full[h.src_sel & 1];
      const CUtensorMap* bmap = &maps.box[h.src_sel & 1][0];
      const int plane_skip = h.src_sel == 3 ? 1 : 0;            // shared-growth layout, buffer B: planes xB | (xA) | x1x2 | x3x4
      for (int set = 0; set < nsets; ++set) {
      const int sb0 = set == 0 ? 0 : nband0, sb1 = set == 0 ? nband0 : nband;
      for (int c = 0; c < nchunk; ++c) {
        // weights of (pass, set, chunk): depend on nobody
        mbar_wait(&sh.wempty[ws], wphase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sh.wfull[ws], kWChunkBytes);   // both halves land here: this CTA's and the peer's multicast
          if (WMC) {
            const int half = static_cast<int>(cluster_ctarank());
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma_load_2d_hint_mc(wring + ws * kWChunkBytes + dx * kWBoxBytes + half * (kWBoxBytes / 2), &maps.wh, &sh.wfull[ws], 0,
                                  h.w_row0 + (dx * nchunk + c) * 3 * COUT + half * (3 * COUT / 2), 3, keep);
          } else {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma_load_2d_hint(wring + ws * kWChunkBytes + dx * kWBoxBytes, &maps.w, &sh.wfull[ws], 0, h.w_row0 + (dx * nchunk + c) * 3 * COUT, keep);
          }
        }
        __syncwarp();
        if (++ws == kWStages) { ws = 0; wphase ^= 1; }
        // activations of this chunk: every CTA must have published the passes that wrote them
        const unsigned need = static_cast<unsigned>(c == 0 ? h.need0 : (c == 1 ? h.need1 : h.need2));
        if (need > known[set] && !(h.dbg & 16384)) {              // 16384: no dependency polling (timing experiments)
          // every lane polls one neighbouring band set (padding lanes: this one).  The spin is RELAXED: ld.acquire.gpu compiles to
          // LDG.STRONG + CCTL.IVALL, and an L1 invalidation per poll made every L1-cached load of the epilogue warps
          // (bias) miss -- ~700 cycles per row.  One acquire after the last poll orders the TMA loads that follow.
          if (ld_relaxed_gpu(my_dep[set]) < need) {
            const long long t0 = clock64();
            while (ld_relaxed_gpu(my_dep[set]) < need) {
              if (clock64() - t0 > NESR_HANG_GUARD_CYCLES) __trap();
            }
          }
          (void)ld_acquire_gpu(my_dep[set]);
          __syncwarp();
          if (!(h.dbg & 8192)) fence_proxy_async_all();          // 8192: no proxy fences (timing experiments)
          known[set] = need;
          if (lane == 0) TS(4, pass);
        }
        const int plane = (c + (c > 0 ? plane_skip : 0)) * plane_px;
        for (int b = sb0; b < sb1; ++b) {
          const BandInfo& bi = sh.band[b];
          const int nrow = bi.rows + 2;
          const bool full_strip = bi.full_strip != 0;
          const uint32_t row_bytes = bi.row_bytes;
          for (int i = 0; i < nrow; ++i) {
            mbar_wait(&sh.empty[stage], phase ^ 1);
            if (elect_one()) {
              if (h.dbg & 4) {
                mbar_arrive(&sh.full[stage]);
              } else {
                mbar_arrive_expect_tx(&sh.full[stage], row_bytes);
                uint8_t* slab = ring + stage * kSlabBytes;
                if (full_strip) {
                  tma_load_2d_hint(slab, amap, &sh.full[stage], 0, plane + bi.op_px[0] + i * bi.op_pitch[0], keep);
                } else {
                  for (int k = 0; k < bi.nop; ++k)
                    tma_load_2d_hint(slab + bi.op_off[k], bmap + bi.op_box[k], &sh.full[stage], 0,
                                     plane + bi.op_px[k] + i * bi.op_pitch[k], keep);
                }
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      }
      h = nh;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // ONE thread runs the whole role.  The tensor pipe queues only ~2 MMAs (tools/commit_probe.cu: of 146 cycles
    // spent away from the issue sequence ~100 are hidden), so the issuer must never be gone for long: no per-row
    // elect / reconvergence, and the barrier polls for the NEXT input row sit between the MMA groups of the current one.
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int ws = 0; uint32_t wphase = 0;
      const uint32_t hw = (__ldg(&passes[0].idesc) >> 7) & 7u;
      const uint32_t hi = umma_desc_hi_sw128();
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(ring));
      const uint32_t w_lo0 = umma_desc_lo(smem_u32(wring));
      constexpr uint32_t kWChunkLo = kWChunkBytes >> 4;
      PassHead h = load_head(passes, 0, npass);
      for (int pass = 0; pass < npass; ++pass) {
        const PassHead nh = load_head(passes, pass + 1, npass);
        const int nchunk = (h.cin + kChunkChannels - 1) / kChunkChannels;
        const uint32_t tparity = static_cast<uint32_t>(pass & 1);
        for (int set = 0; set < nsets; ++set) {
        const int sb0 = set == 0 ? 0 : nband0, sb1 = set == 0 ? nband0 : nband;
        for (int c = 0; c < nchunk; ++c) {
          const int rem = (h.cin - c * kChunkChannels) >> 4;
          const int ks = rem < 4 ? rem : 4;
          const bool last_chunk = c + 1 == nchunk;
          const bool first_chunk = c == 0;
          mbar_wait(&sh.wfull[ws], wphase);
          const uint32_t w_lo = w_lo0 + ws * kWChunkLo;
          const bool mma_on = !(h.dbg & 2);
          const int variant = (ks == 4 ? 0 : 4) + (first_chunk ? 2 : 0) + (last_chunk ? 1 : 0);
          for (int b = sb0; b < sb1; ++b) {
            const int rows = sh.band[b].rows, slot0 = sh.band[b].slot0;
            switch (variant) {                                  // trunk passes only have 4- and 2-k-step chunks
              case 0: sweep_band<4, false, false>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              case 1: sweep_band<4, false, true>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              case 2: sweep_band<4, true, false>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              case 3: sweep_band<4, true, true>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              case 4: sweep_band<2, false, false>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              case 5: sweep_band<2, false, true>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              case 6: sweep_band<2, true, false>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
              default: sweep_band<2, true, true>(sh, rows, slot0, tmem_base, hw, hi, a_lo0, w_lo, tparity, stage, phase, mma_on); break;
            }
            if (last_chunk && b == sb0) TS(5, pass);
          }
          if (WMC) umma_commit_mc(&sh.wempty[ws], 3); else umma_commit(&sh.wempty[ws]);
          if (++ws == kWStages) { ws = 0; wphase ^= 1; }
          if (last_chunk) TS(0, pass);
        }
        }
        h = nh;
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;                          // rows alternate between the two epilogue groups
    const int m = quarter * 32 + lane;                          // TMEM lane == MMA row == pixel
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int pass = 0; pass < npass; ++pass) {
      // The epilogue warps are instruction-latency bound (one or two warps per scheduler, ~4 cycles per dependent
      // instruction): the generic epilogue16() path cost ~300 instructions = 1200 cycles per row.  A trunk pass is one of
      // two kinds, fixed for the whole pass, so the row loop below is straight-line code specialised at pass level:
      //   conv1..4 : v = lrelu(acc + bias)                               -> 16-bit, channels [coff, coff+32) of this buffer
      //   conv5    : v = (acc + bias)*0.2 + trunk [; v = v*0.2 + rrdb_in] -> fp32 trunk [+ rrdb], 16-bit x of the next block
      const ConvParams* pp = passes + pass;
      const int dbg = __ldg(&pp->debug_flags);
      const float* res1 = pp->res1;
      const float* res2 = pp->res2;
      float* dst32a = pp->dst32a;
      float* dst32b = pp->dst32b;
      const float s1 = pp->s1, s2 = pp->s2;
      const int c_off = pp->c_off, fmt16 = pp->dst16_fmt, lrelu = pp->lrelu;
      const int coff16 = pp->dst16_coff;
      uint16_t* const base16 = reinterpret_cast<uint16_t*>(pp->dst16) + static_cast<size_t>(coff16 >> 6) * pp->dst16_plane_px * 64 + (coff16 & 63);
      float bias_r[COUT];                                       // once per pass, in registers
      {
        const float4* b4 = reinterpret_cast<const float4*>(pp->bias);
#pragma unroll
        for (int k = 0; k < COUT / 4; ++k) {
          const float4 bv = __ldg(b4 + k);
          bias_r[4 * k] = bv.x; bias_r[4 * k + 1] = bv.y; bias_r[4 * k + 2] = bv.z; bias_r[4 * k + 3] = bv.w;
        }
      }
      const uint32_t tparity = static_cast<uint32_t>(pass & 1);
      for (int set = 0; set < nsets; ++set) {
      const int sb0 = set == 0 ? 0 : nband0, sb1 = set == 0 ? nband0 : nband;
      for (int b = sb0; b < sb1; ++b) {
        const int rows = sh.band[b].rows, slot0 = sh.band[b].slot0;
        const int px0 = sh.lane_px[b][m], pitch = sh.lane_pitch[b][m], my_rows = sh.lane_rows[b][m];
        const bool band_on = px0 >= 0 && !(dbg & 1);
        for (int j = 0; j < rows; ++j) {
          const int slot = slot0 + j;
          if ((slot & 1) != group) continue;
          const bool lane_on = band_on && j < my_rows;
          const int P = px0 + j * pitch;
          // blocked fp32 trunk layout: [pixel/32][ch/8][pixel%32][ch%8]; this lane's 32 channels are 4 runs of 8 floats
          const size_t toff = (static_cast<size_t>(P >> 5) * 8 + (c_off >> 3)) * 256 + (static_cast<size_t>(P & 31) << 3);
          // residual rows are fetched BEFORE waiting for the accumulator: their latency hides behind the MMAs
          float r1[COUT], r2[COUT];
          if (lane_on && res1) {
            // a tile group's fp32 trunk (35 MB) fits in L2 next to its dense-block planes: plain accesses (1 % faster than
            // evict_first streaming, which is what the whole-frame kernel needs); the RRDB input, read once in three
            // blocks, keeps streaming
#pragma unroll
            for (int q = 0; q < COUT / 8; ++q) ldg256(res1 + toff + q * 256, &r1[q * 8]);
          }
          if (lane_on && res2) {
#pragma unroll
            for (int q = 0; q < COUT / 8; ++q) ldg256_stream(res2 + toff + q * 256, &r2[q * 8]);
          }
          EPI_T(et0);
          mbar_wait(&sh.tfull[slot], tparity);
          EPI_T(et1);
          tc_fence_after();
          __syncwarp();
          const uint32_t taddr = lane_base + static_cast<uint32_t>(slot) * COUT;
          uint32_t r[COUT / 16][16];
#pragma unroll
          for (int c = 0; c < COUT / 16; ++c) tmem_ld16(taddr + c * 16, r[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < COUT / 16; ++c) tmem_st16_zero(taddr + c * 16);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&sh.tempty[slot]);
          EPI_T(et2);
          if (lane_on) {
            float v[COUT];
#pragma unroll
            for (int k = 0; k < COUT; ++k) v[k] = __uint_as_float(r[k >> 4][k & 15]) + bias_r[k];
            if (!res1) {
              if (lrelu) {
#pragma unroll
                for (int k = 0; k < COUT; ++k) v[k] = fmaxf(v[k], 0.2f * v[k]);     // LeakyReLU(0.2): slope < 1
              }
            } else {
#pragma unroll
              for (int k = 0; k < COUT; ++k) v[k] = fmaf(v[k], s1, r1[k]);
              if (res2) {
#pragma unroll
                for (int k = 0; k < COUT; ++k) v[k] = fmaf(v[k], s2, r2[k]);
              }
              if (!(dbg & 4096)) {
#pragma unroll
                for (int q = 0; q < COUT / 8; ++q) stg256f(dst32a + toff + q * 256, &v[q * 8]);
                if (dst32b) {
#pragma unroll
                  for (int q = 0; q < COUT / 8; ++q) stg256f_stream(dst32b + toff + q * 256, &v[q * 8]);
                }
              }
            }
            if (!(dbg & 16)) {
              uint32_t w[COUT / 2];
              if (fmt16) {
#pragma unroll
                for (int k = 0; k < COUT / 2; ++k) w[k] = pack2(v[2 * k], v[2 * k + 1], 1);
              } else {
#pragma unroll
                for (int k = 0; k < COUT / 2; ++k) w[k] = pack2(v[2 * k], v[2 * k + 1], 0);
              }
              uint16_t* dst = base16 + static_cast<size_t>(P) * 64;
              stg256(dst, reinterpret_cast<const uint32_t(&)[8]>(w[0]));
              stg256(dst + 16, reinterpret_cast<const uint32_t(&)[8]>(w[8]));
            }
          }
#if NESR_PROF
          { const long long et3 = clock64(); EPI_ACC(0, pass, et1 - et0); EPI_ACC(1, pass, et2 - et1); EPI_ACC(2, pass, et3 - et2); EPI_ACC(3, pass, 1); }
#endif
        }
      }
      // publish the pass: generic-proxy stores -> TMA (async proxy) reads of any CTA
      if (threadIdx.x == 64) TS(1, pass);
      if (__ldg(&pp->trunk_no_publish)) continue;               // (uniform) covered by the next pass's publish
      if (!(dbg & 8192)) fence_proxy_async_all();
      if (!(dbg & 32768)) epi_bar_sync();                       // 32768: no named barrier / release (timing experiments)
      if (threadIdx.x == 64) {
        TS(2, pass);
        if (!(dbg & 32768)) st_release_gpu(prog + static_cast<size_t>(blockIdx.x * 2 + set) * kProgStride, static_cast<unsigned>(pass + 1));
        TS(3, pass);
      }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (WMC) cluster_sync_all();                                  // no CTA leaves while its peer may still multicast into it
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#if NESR_PROF
  if ((passes[0].debug_flags & 1024) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 40)) {
    const long long t0 = sh.ts[5][0];
    for (int q = 0; q < kTracePasses && q < npass; ++q)
      printf("[trunk blk %d pass %d cin=%d] acquired %lld  first_full_last_chunk %lld  mma_issued %lld  epi_rows_done %lld  epi_synced %lld  published %lld\n",
             (int)blockIdx.x, q, passes[q].cin, sh.ts[4][q] - t0, sh.ts[5][q] - t0, sh.ts[0][q] - t0, sh.ts[1][q] - t0, sh.ts[2][q] - t0, sh.ts[3][q] - t0);
    for (int q = 0; q < kTracePasses && q < npass; ++q)
      printf("[trunk epi blk %d pass %d cin=%d] rows %lld  wait_tfull %lld  tmem_ld_zero_arrive %lld  math_stores %lld\n", (int)blockIdx.x, q,
             passes[q].cin, sh.epi_acc[3][q], sh.epi_acc[0][q], sh.epi_acc[1][q], sh.epi_acc[2][q]);
  }
#endif
}

}  // namespace

cudaError_t conv3x3_trunk_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv3x3_trunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv3x3_trunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t launch_conv3x3_trunk(const TrunkMaps& maps, const ConvParams* d_passes, int npass, unsigned* d_gbar, int grid,
                                 cudaStream_t stream, bool weight_multicast) {
  if (grid <= 0 || npass <= 0) return cudaSuccess;
  if (grid > 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaMemsetAsync(d_gbar, 0, static_cast<size_t>(2 * grid) * kProgStride * sizeof(unsigned), stream);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;                 // co-residency guarantee for the arrival counter
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (weight_multicast && (grid & 1) == 0) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, conv3x3_trunk_kernel<true>, maps, d_passes, npass, d_gbar);
  }
  return cudaLaunchKernelEx(&cfg, conv3x3_trunk_kernel<false>, maps, d_passes, npass, d_gbar);
}

}  // namespace nesr
