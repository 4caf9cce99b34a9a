// conv3x3_trunk.cu -- the 23 x 3 residual dense blocks of one L2-resident tile group as ONE persistent kernel,
// organised so that consecutive layer passes OVERLAP instead of draining the SM between them.
//
// What the earlier persistent kernel (conv3x3_body.cu) taught (profiles/r1_body_vs_trunk_dram.txt, r1 launch list):
//   * swept over a whole 1080p frame, every layer pass re-reads its 64..192 input channels from HBM -- the
//     frame's dense-block buffer (213 MB) cannot live in the 126 MB L2, and the trunk runs at the HBM roofline;
//   * shrinking the working set made it SLOWER: every pass costs ~10 us of fixed latency (drain, grid-wide arrival
//     counter, reload up to 110 KB of weights, refill the ring) and small batches have short passes.
// This kernel removes the fixed cost, so that tile groups small enough for L2 pay off:
//
//   TMEM-RESIDENT BANDS.  A CTA owns at most 8 output rows (a few bands of 128-pixel strips); a row's TMEM slot is 64
//   fp32 columns -- half A | half B -- and all 8 x 64 = 512 columns stay resident while a dense block is computed.
//   CHUNK-MAJOR, MERGED SWEEPS (round 2).  A sweep streams the band's input rows of ONE 64-channel plane once and
//   accumulates into everything that plane feeds next: the kernel is bound by the bytes it pulls through L2 (TMA alone
//   needed 75 % of the round-1 kernel's time; profiles/r1_trunk_experiments.txt) and by N = 96 MMAs running at 85 % of the
//   tensor rate, and merging attacks both.  Round 1 ran conv1..conv5 as six passes = 14 plane sweeps of N = 96 per dense
//   block; now a block is EIGHT sweeps (layout.h TrunkSweep), six of them N = 192 (99 % of the tensor rate):
//       S1 x -> conv1|conv2   S2 x1 -> conv2   S3 x -> conv3|conv4   S4 x1,x2 -> conv3|conv4   S5 x3 -> conv4
//       S6 x -> conv5         S7 x1,x2 -> conv5                      S8 x3,x4 -> conv5
//   conv1/conv3/conv5[0:32] accumulate in half A of a slot, conv2/conv4/conv5[32:64] in half B; the two single-layer
//   sweeps (S2, S5) run as N = 160 MMAs whose weight rows over half A are zero (three N = 32 MMAs would cost 162 cycles
//   against 80).  The epilogue still drains 32 columns at a time, in six passes per block, each as soon as its half is complete.
//   The weights stream too: one [3 dx][192|160][64] box set (72 KB) per sweep through a 2-deep ring.
//   DEPENDENCIES ARRIVE LAST.  The sweep that needs the newest channels (S2, S4, S5, S8 and the next block's S1) is the one
//   that completes an accumulator; the sweeps between them (S3, S6, S7) need nothing new.
//   ROW-GRANULAR HAND-OVER (round 2).  The round-1 kernel published "pass complete" once per pass and CTA, so every pass
//   was a serial chain  acquire -> first slab -> dependent sweep -> epilogue tail -> fence + barrier + publish  (22k cycles
//   per conv1-4 pass against 9-22k cycles of MMA work: profiles/r1_trunk_chain_trace.txt).  Now every CTA publishes a
//   ROW counter (pass * 32 + rows stored) in its own 128-byte line, advanced by a dedicated PUBLISHER warp as the epilogue
//   warps signal stored rows on shared-memory mbarriers, and the TMA producer waits per slab row for exactly the rows of
//   the neighbouring CTAs (and of its own) that the slab covers -- a host-built table [slab row][dependency] of row counts.
//   The gpu-scope release (a membar) is paid by a warp that has nothing else to do, and nothing in the kernel is a
//   CTA-wide barrier any more.
//   PHASE-ORDERED ROWS (round 2, layout.h trunk_order).  Swept top to bottom, a dependent sweep's first slab row is the last
//   row the CTA above completes: no sweep could start before its neighbours' previous pass had ended.  Every CTA now sweeps
//   its input rows (halo rows included, bands interleaved) sorted by the tile row's residue mod 8 in the sequence
//   3 4 2 5 1 6 0 7 -- which is also the order in which its output rows complete, in every CTA of the group whatever its row
//   offset -- so what a sweep needs first is what the previous pass finished first, everywhere, and the hand-over
//   (drain -> store -> publish -> acquire -> TMA, ~10k cycles) hides behind the remaining 8 rows of the sweep.
//   CONV5'S RESIDUALS LIVE IN THE ACCUMULATOR (round 2).  The conv3 / conv4 epilogues leave  bias + res1/s1 [+ res2/(s1 s2)]
//   of conv5's 32 channels in the half they have drained instead of zeros, so conv5's epilogue is a multiply and stores: its
//   residual loads were what the next block's first sweep waited for.
//
// Memory-model argument for the publisher (PTX causality order is transitive): epilogue thread's st.global (+ its
// fence.proxy.async) -> mbarrier.arrive (release.cta) -> publisher's mbarrier wait (acquire.cta) -> publisher's
// st.release.gpu -> consumer lane's ld.acquire.gpu -> __syncwarp -> fence.proxy.async -> TMA load.  Round 1 used the same
// shape with bar.sync in place of the mbarrier.
//
// Write-after-read safety of the single-buffered growth planes (x itself ping-pongs between two buffers).  A CTA overwrites
// pixel p of plane x_k (block n+1) when its conv_k epilogue drains p's row, i.e. after a sweep of block n+1 has swept the
// three input rows around p -- rows that hold, for every CTA C that reads p (C owns a pixel within one row and one column of
// p), one of C's own pixels of a plane C stored in block n+1 or in conv5 of block n.  C stored that pixel from an epilogue
// that ran after a commit of one of ITS block-n+1 sweeps (or of S8 of block n), the commit covers every MMA C issued before
// it, and C issues sweeps in program order: every sweep of block n that reads x_k (for x3 / x4: S8, the last one) had
// consumed its slabs before p is overwritten.  The argument uses the order of SWEEPS, not of rows inside a sweep, so it
// holds for any row order.
//
// Roles (fold_roles.cuh explains the row fold itself): warp 0 TMA producer (weights + row slabs), warp 1 MMA issuer,
// warps 2..9 epilogue (two groups alternating rows), warp 10 publisher.  Launched cooperatively (grid <= #SMs).
#include <stdio.h>

#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace nesr {

namespace {

constexpr int COUT = 32;                               // channels one epilogue pass drains (half a TMEM row slot)
constexpr int kSlotCols = 64;                          // fp32 columns of a TMEM row slot: half A | half B
constexpr int kThreads = 352;
constexpr int kSlabPx = 136;
constexpr int kSlabBytes = kSlabPx * 128;              // 17408
constexpr int kStages = 4;                             // activation slab ring
constexpr int kWStages = 2;                            // sweep weight ring
constexpr int kWBoxRows = 192;                         // rows of one dx box (160 for half-B single sweeps)
constexpr int kWStageBytes = 3 * kWBoxRows * 128;      // three dx boxes = 73728
constexpr int kSlots = kTrunkMaxRows;                  // TMEM row slots (512 / 64)
constexpr int kMaxBands = kTrunkMaxBands;
constexpr int kProgUnit = 32;                          // progress value = pass * kProgUnit + rows stored (rows <= 8)
#if NESR_PROF
constexpr int kTrace = 48;
// sweeps [kTrace0 * 8, +kTrace) and passes [kTrace0 * 6, +kTrace) are recorded: dense block kTrace0 onwards (0 when the net is shorter)
#define TS(k, idx) do { const int i_ = (idx) - ts_base[(k) == 0 || (k) == 4 ? 0 : 1]; if (i_ >= 0 && i_ < kTrace) sh.ts[k][i_] = clock64(); } while (0)
#else
#define TS(k, idx) do {} while (0)
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
  uint64_t tfull[2][kSlots], tempty[2][kSlots];        // per TMEM half (A, B) and row slot
  uint64_t stored[kSlots];                             // epilogue -> publisher: this row's stores are done (one arrive per warp)
  uint32_t tmem_slot;
  int32_t nband, nrows;                                // bands / output rows of this CTA
  BandInfo band[kMaxBands];
  TrunkOrder order;                                    // phase order of the CTA's input rows, completion order of its output rows (layout.h)
  // what the MMA issuer does with the t-th slab row of that order (the same in every sweep): first slot it feeds | slots - 1 << 4 |
  // weight blocks skipped at the band's upper end << 6 | slots it touches first << 8 | slots it completes << 16
  uint32_t step[kTrunkMaxSlabRows];
  int32_t lane_px[kMaxBands][128];                     // flat pixel of (r0, x) of each MMA lane, or -1 (masked lane)
  int32_t lane_pitch[kMaxBands][128];
  int32_t lane_rows[kMaxBands][128];                   // band rows [0, lane_rows) belong to the lane's piece
  uint8_t need_rows[kTrunkMaxSlabRows][kTrunkMaxDeps]; // [slab row of the CTA][dependency lane]: rows that CTA must have stored
  alignas(16) float bias[8][2 * COUT];                             // per epilogue warp: bias of the pass (both halves of a paired conv5 pass)
#if NESR_PROF
  long long ts[5][kTrace];                             // time stamps of the dependency chain (debug_flags & 1024)
#endif
};

constexpr int kRingBytes = kWStages * kWStageBytes + kStages * kSlabBytes;      // 147456 + 69632
constexpr int kSmemBytes = kRingBytes + static_cast<int>(sizeof(Shared)) + 1024;
static_assert(kSmemBytes <= 232448, "trunk kernel: shared memory budget");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* ptr, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
constexpr int kProgStride = 32;                        // one 128-byte line per CTA
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ TrunkSweep load_sweep(const TrunkSweep* sweeps, int i, int n) {
  TrunkSweep s{};
  if (i < n) {
    const int4* p = reinterpret_cast<const int4*>(sweeps + i);
    const int4 a = __ldg(p), b = __ldg(p + 1);
    s.src_sel = a.x; s.plane = a.y; s.w_row0 = a.z; s.nb = a.w; s.ks = b.x; s.need = b.y; s.flags = b.z;
  }
  return s;
}

// One sweep over the CTA's input rows, executed by the single MMA-issuing thread.  KS k-steps per (row, dx).  SINGLE: an N = 160
// sweep into half B whose weight rows over half A are zero (D starts at column 32 of the first slot); otherwise both halves of
// every slot, N = 192.  Slab row i of a band (input row i-1) feeds output rows i-2, i-1, i = three consecutive slots; at the
// band's ends the MMA is clamped to the rows that exist (N and the first weight row shrink).  The rows are taken in the CTA's
// PHASE ORDER (layout.h trunk_order), bands interleaved.  wait[X]: this sweep is the first toucher of half X after its drain:
// every slot is waited for (drained + re-zeroed) before the first MMA that touches it; commit[X]: this sweep completes half X:
// every output row is committed to the epilogue as soon as its third contribution is issued.
// HALF_A (the two single-layer sweeps S2 / S5, KS = 2): the plane's first 32 channels only, loaded as a half-width slab -- 64-byte rows,
// SWIZZLE_64B (8.7 KB instead of 17 KB per slab row: these sweeps were bound by the slab stream, 8.7k cycles against 4.8k of MMA work);
// the dx shift is then a 64-byte offset of the A descriptor.
template <int KS, bool SINGLE, bool HALF_A = false>
__device__ __forceinline__ void sweep_cta(Shared& sh, const uint32_t tmem_base, const uint32_t hw,
                                          const uint32_t hi, const uint32_t a_lo0, const uint32_t w_lo, const uint32_t box_lo,
                                          const bool wait_a, const uint32_t par_a, const bool wait_b, const uint32_t par_b,
                                          const bool commit_a, const bool commit_b, int& stage, uint32_t& phase, const bool mma_on) {
  constexpr uint32_t kSlabLo = kSlabBytes >> 4;
  constexpr uint32_t kBlkLo = (kSlotCols * 128) >> 4;           // weight rows of one output row (64 rows of 128 bytes)
  constexpr uint32_t kDx = HALF_A ? 4u : 8u;                    // one pixel row of the slab in descriptor units (16 bytes)
  const int n_in = sh.order.n_in;
  if (n_in == 0) return;                                        // a CTA the deal left without rows
  const uint32_t id1 = umma_idesc_f16(hw, SINGLE ? 32u : 64u), id2 = umma_idesc_f16(hw, SINGLE ? 96u : 128u),
                 id3 = umma_idesc_f16(hw, SINGLE ? 160u : 192u);
  const bool waits = wait_a || wait_b, commits = commit_a || commit_b;
  auto wait_slots = [&](uint32_t m) {                           // drained + re-zeroed?
    while (m) {
      const int sl = __ffs(m) - 1;
      m &= m - 1;
      if (wait_a) mbar_wait(&sh.tempty[0][sl], par_a);
      if (wait_b) mbar_wait(&sh.tempty[1][sl], par_b);
    }
  };
  uint32_t w = sh.step[0];
  if (waits) wait_slots((w >> 8) & 0xffu);
  mbar_wait(&sh.full[stage], phase);
  tc_fence_after();
  for (int t = 0; t < n_in; ++t) {
    const uint32_t nsl = (w >> 4) & 3u;
    const uint32_t d = tmem_base + (w & 15u) * kSlotCols + (SINGLE ? 32u : 0u);
    const uint32_t id = nsl == 2u ? id3 : (nsl == 1u ? id2 : id1);
    const uint32_t b_lo = w_lo + ((w >> 6) & 3u) * kBlkLo;
    const uint32_t a_lo = a_lo0 + stage * kSlabLo;
    if (mma_on) {
      if (HALF_A) umma_f16_2ksteps_half_a(d, a_lo, b_lo, hi, id, umma_desc_hi_sw64());
      else umma_f16_ksteps<KS>(d, a_lo, b_lo, hi, id);
    }
    // while those run: are the next input row and the slots it touches first ready?
    const int nstage = stage + 1 == kStages ? 0 : stage + 1;
    uint32_t wn = 0;
    if (t + 1 < n_in) {
      wn = sh.step[t + 1];
      if (waits) wait_slots((wn >> 8) & 0xffu);
      mbar_wait(&sh.full[nstage], nstage == 0 ? phase ^ 1 : phase);
      tc_fence_after();
    }
    if (mma_on) {
      if (HALF_A) {
        umma_f16_2ksteps_half_a(d, a_lo + kDx, b_lo + box_lo, hi, id, umma_desc_hi_sw64());
        umma_f16_2ksteps_half_a(d, a_lo + 2 * kDx, b_lo + 2 * box_lo, hi, id, umma_desc_hi_sw64());
      } else {
        umma_f16_ksteps<KS>(d, a_lo + kDx, b_lo + box_lo, hi, id);
        umma_f16_ksteps<KS>(d, a_lo + 2 * kDx, b_lo + 2 * box_lo, hi, id);
      }
    }
    umma_commit(&sh.empty[stage]);                              // slab may be overwritten once these MMAs have read it
    if (commits) {
      uint32_t m = (w >> 16) & 0xffu;                           // output rows that now have their three contributions
      while (m) {
        const int sl = __ffs(m) - 1;
        m &= m - 1;
        if (commit_a) umma_commit(&sh.tfull[0][sl]);
        if (commit_b) umma_commit(&sh.tfull[1][sl]);
      }
    }
    w = wn;
    stage = nstage;
    if (stage == 0) phase ^= 1;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_trunk_kernel(const __grid_constant__ TrunkMaps maps, const ConvParams* __restrict__ passes, const int npass,
                     const TrunkSweep* __restrict__ sweeps, const int nsweep, unsigned* __restrict__ prog) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wring = smem;
  uint8_t* ring = smem + kWStages * kWStageBytes;
  Shared& sh = *reinterpret_cast<Shared*>(smem + kRingBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup: barriers, TMEM, band geometry (identical for every pass of the trunk) ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.full[0]); tma_prefetch_desc(&maps.full[1]);
    tma_prefetch_desc(&maps.w192); tma_prefetch_desc(&maps.w160);
    tma_prefetch_desc(&maps.half[0]); tma_prefetch_desc(&maps.half[1]);
    for (int i = 0; i < kWStages; ++i) { mbar_init(&sh.wfull[i], 1); mbar_init(&sh.wempty[i], 1); }
    for (int i = 0; i < kStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&sh.tfull[0][i], 1); mbar_init(&sh.tfull[1][i], 1);
      mbar_init(&sh.tempty[0][i], 128); mbar_init(&sh.tempty[1][i], 128);
      mbar_init(&sh.stored[i], 4);
    }
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
      sh.nrows = slot0;
      int brows[kMaxBands], by0[kMaxBands];
      for (int b = 0; b < nband; ++b) {
        const FoldBand band = p0.bands[band_begin + b];
        brows[b] = band.rows; by0[b] = band.r0;
        for (int sgi = 0; sgi < band.nseg; ++sgi) {
          const FoldSeg sg = p0.segs[band.seg0 + sgi];
          if (band.r0 < sg.h) { by0[b] = sg.y0 + band.r0; break; }   // pieces of a packed strip start at multiples of 8 tile rows
        }
      }
      trunk_order(brows, by0, nband, sh.order);
      uint32_t swept[kMaxBands] = {0, 0, 0, 0}, touched = 0;
      for (int t = 0; t < sh.order.n_in; ++t) {
        const int b = sh.order.in_band[t], i = sh.order.in_row[t];
        const int rows = sh.band[b].rows, s0 = sh.band[b].slot0;
        const int lo = i - 2 < 0 ? 0 : i - 2, hi_row = i > rows - 1 ? rows - 1 : i;
        uint32_t first = 0, complete = 0;
        swept[b] |= 1u << i;
        for (int j = lo; j <= hi_row; ++j) {
          if (!((touched >> (s0 + j)) & 1u)) first |= 1u << (s0 + j);
          if (((swept[b] >> j) & 7u) == 7u) complete |= 1u << (s0 + j);   // slab rows j, j + 1, j + 2
        }
        touched |= first;
        sh.step[t] = static_cast<uint32_t>(s0 + lo) | static_cast<uint32_t>(hi_row - lo) << 4 | static_cast<uint32_t>(lo - (i - 2)) << 6 |
                     first << 8 | complete << 16;
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
    } else if (threadIdx.x < 128 + kTrunkMaxSlabRows * kTrunkMaxDeps / 4) {   // the CTA's row-dependency table, 4 bytes per thread
      const int i = threadIdx.x - 128;
      reinterpret_cast<uint32_t*>(&sh.need_rows[0][0])[i] =
          __ldg(reinterpret_cast<const uint32_t*>(p0.trunk_need + static_cast<size_t>(blockIdx.x) * kTrunkMaxSlabRows * kTrunkMaxDeps) + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  if (warp >= 2 && warp < 6) {                                  // every MMA accumulates: start from zero
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = 0; c < 512; c += 16) tmem_st16_zero(t0 + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int nband = sh.nband;
  [[maybe_unused]] const int dbg0 = NESR_PROF ? __ldg(&passes[0].debug_flags) : 0;
#if NESR_PROF
  const int trace_blk = nsweep >= 40 * kSweepsPerBlock ? 30 : 0;
  const int ts_base[2] = {trace_blk * kSweepsPerBlock, trace_blk * 6};
#endif

  if (warp == 0) {
    // =========================== TMA producer ===========================
    int stage = 0; uint32_t phase = 0;                          // slab ring
    int ws = 0; uint32_t wphase = 0;                            // weight ring
    unsigned known = 0;                                         // this CTA's row requirements on passes [0, known) are all met
    unsigned seen = 0;                                          // last value read from this lane's dependency word
    const int plane_px = __ldg(&passes[0].src_plane_px);
    const uint64_t keep = l2_policy_evict_last();               // dense-block activations and weights: stay in L2
    // one lane per dependency: the progress word of a CTA owning pixels of this CTA's slab rows (lane 0: its own; padding: its own)
    const unsigned* my_dep = prog + static_cast<size_t>(__ldg(passes[0].trunk_deps + blockIdx.x * kTrunkMaxDeps + lane)) * kProgStride;
    TrunkSweep sw = load_sweep(sweeps, 0, nsweep);
    for (int si = 0; si < nsweep; ++si) {
      const TrunkSweep nsw = load_sweep(sweeps, si + 1, nsweep);  // next sweep, fetched early
      // weights of the sweep: depend on nobody
      mbar_wait(&sh.wempty[ws], wphase ^ 1);
      if (elect_one()) {
        const uint32_t box_bytes = static_cast<uint32_t>(sw.nb) * 128u;
        mbar_arrive_expect_tx(&sh.wfull[ws], 3u * box_bytes);
        const CUtensorMap* wmap = sw.nb == kWBoxRows ? &maps.w192 : &maps.w160;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          tma_load_2d_hint(wring + ws * kWStageBytes + dx * box_bytes, wmap, &sh.wfull[ws], 0, sw.w_row0 + dx * sw.nb, keep);
      }
      __syncwarp();
      if (++ws == kWStages) { ws = 0; wphase ^= 1; }
      // activations.  A plane whose newest channels were written by epilogue pass need-1 is requested slab row by slab row,
      // each as soon as the rows it covers have been stored everywhere (host-built table need_rows).
      const unsigned need = static_cast<unsigned>(sw.need);
      const bool rowwise = need > known && !(dbg0 & 16384);     // 16384: no dependency waits (timing experiments)
      const unsigned row_base = (need - 1) * kProgUnit;
      const bool half = sw.ks == 2 && (sw.flags & kSweepSingleB);   // S2 / S5: the plane's first 32 channels, half-width slab rows
      const CUtensorMap* amap = half ? &maps.half[sw.src_sel & 1] : &maps.full[sw.src_sel & 1];
      const CUtensorMap* bmap = half ? &maps.hbox[sw.src_sel & 1][0] : &maps.box[sw.src_sel & 1][0];
      const int plane = sw.plane * plane_px;
      const int n_in = sh.order.n_in;
      for (int t = 0; t < n_in; ++t) {                          // slab rows of the CTA in phase order
        {
          const BandInfo& bi = sh.band[sh.order.in_band[t]];
          const int i = sh.order.in_row[t];
          const bool full_strip = bi.full_strip != 0;
          const uint32_t row_bytes = half ? bi.row_bytes >> 1 : bi.row_bytes;
          if (rowwise) {
            const unsigned req = sh.need_rows[t][lane];
            unsigned target = req ? row_base + req : 0u;
            if ((dbg0 & 262144) && req) target = need * kProgUnit;   // 262144: wait for whole passes (timing experiments)
            if (__any_sync(0xffffffffu, seen < target)) {
              // Somebody has to look: EVERY lane refreshes its word (one L2 round trip for the warp either way), so that rows
              // further down -- which usually depend on CTAs this row does not -- find their requirement already seen.
              // Every poll is an acquire (LDG.STRONG + CCTL.IVALL): nothing in this kernel lives in L1 (bias is in registers,
              // the trunk rows stream), and a separate acquire after a relaxed spin costs one more round trip.  The generic ->
              // async proxy fence of the chain is executed by the PUBLISHER before its release (a proxy fence in this thread
              // waits for the TMA loads in flight: measured ~1k cycles each, 8 % of the kernel).
              const long long t0 = clock64();
              do {
                seen = ld_acquire_gpu(my_dep);
                if (clock64() - t0 > NESR_HANG_GUARD_CYCLES) __trap();
              } while (seen < target);
              __syncwarp();
            }
          }
          if (lane == 0 && t == 0) TS(4, si);
          mbar_wait(&sh.empty[stage], phase ^ 1);
          if (elect_one()) {
            if (dbg0 & 4) {
              mbar_arrive(&sh.full[stage]);
            } else {
              mbar_arrive_expect_tx(&sh.full[stage], row_bytes);
              uint8_t* slab = ring + stage * kSlabBytes;
              if (full_strip) {
                tma_load_2d_hint(slab, amap, &sh.full[stage], 0, plane + bi.op_px[0] + i * bi.op_pitch[0], keep);
              } else {
                for (int k = 0; k < bi.nop; ++k)
                  tma_load_2d_hint(slab + (half ? bi.op_off[k] >> 1 : bi.op_off[k]), bmap + bi.op_box[k], &sh.full[stage], 0,
                                   plane + bi.op_px[k] + i * bi.op_pitch[k], keep);
              }
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (rowwise) known = need;
      sw = nsw;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // ONE thread runs the whole role.  The tensor pipe queues only ~2 MMAs (tools/commit_probe.cu), so the issuer must
    // never be gone for long: no per-row elect / reconvergence, and the barrier polls for the NEXT input row sit between
    // the MMA groups of the current one.
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int ws = 0; uint32_t wphase = 0;
      uint32_t nwait_a = 0, nwait_b = 0;                        // waits done so far on the drained-slot barriers of each half
      const uint32_t hw = (__ldg(&passes[0].idesc) >> 7) & 7u;
      const uint32_t hi = umma_desc_hi_sw128();
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(ring));
      const uint32_t w_lo0 = umma_desc_lo(smem_u32(wring));
      constexpr uint32_t kWStageLo = kWStageBytes >> 4;
      const bool mma_on = !(dbg0 & 2);
      TrunkSweep sw = load_sweep(sweeps, 0, nsweep);
      for (int si = 0; si < nsweep; ++si) {
        const TrunkSweep nsw = load_sweep(sweeps, si + 1, nsweep);
        const bool wait_a = sw.flags & kSweepWaitA, wait_b = sw.flags & kSweepWaitB;
        const bool commit_a = sw.flags & kSweepCommitA, commit_b = sw.flags & kSweepCommitB;
        const uint32_t par_a = (nwait_a & 1u) ^ 1u, par_b = (nwait_b & 1u) ^ 1u;   // the k-th wait passes once k drains have happened
        const uint32_t box_lo = (static_cast<uint32_t>(sw.nb) * 128u) >> 4;
        mbar_wait(&sh.wfull[ws], wphase);
        const uint32_t w_lo = w_lo0 + ws * kWStageLo;
        const int variant = ((sw.flags & kSweepSingleB) ? 2 : 0) + (sw.ks == 4 ? 0 : 1);
        switch (variant) {
          case 0: sweep_cta<4, false>(sh, tmem_base, hw, hi, a_lo0, w_lo, box_lo, wait_a, par_a, wait_b, par_b, commit_a, commit_b, stage, phase, mma_on); break;
          case 1: sweep_cta<2, false>(sh, tmem_base, hw, hi, a_lo0, w_lo, box_lo, wait_a, par_a, wait_b, par_b, commit_a, commit_b, stage, phase, mma_on); break;
          case 2: sweep_cta<4, true>(sh, tmem_base, hw, hi, a_lo0, w_lo, box_lo, wait_a, par_a, wait_b, par_b, commit_a, commit_b, stage, phase, mma_on); break;
          default: sweep_cta<2, true, true>(sh, tmem_base, hw, hi, a_lo0, w_lo, box_lo, wait_a, par_a, wait_b, par_b, commit_a, commit_b, stage, phase, mma_on); break;
        }
        umma_commit(&sh.wempty[ws]);
        if (++ws == kWStages) { ws = 0; wphase ^= 1; }
        nwait_a += wait_a; nwait_b += wait_b;
        TS(0, si);
        sw = nsw;
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // =========================== epilogue (warps 2..9) ===========================
    // The epilogue warps are latency bound (one or two warps per scheduler; tcgen05.ld / st waits, L2 round trips), so the
    // two kinds of pass are separate straight-line code paths and everything that does not change between passes -- which
    // rows this warp drains, where their pixels live -- is computed once:
    //   conv1..4 : v = lrelu(acc + bias)                               -> 16-bit, channels [coff, coff+32) of this buffer
    //   conv5    : v = (acc + bias)*0.2 + trunk [; v = v*0.2 + rrdb_in] -> fp32 trunk [+ rrdb], 16-bit x of the next block
    // conv5's two 32-channel passes complete together (sweep S8 commits both halves of a row at once) and are drained
    // TOGETHER, row by row: drained one after the other, the second pass's rows all waited behind the first's and the
    // block's last sweep was followed by a 20k-cycle epilogue tail before the next block could start.  Its fp32 trunk rows
    // are prefetched one row ahead (both halves, two register sets): issued just before they were needed, their L2 round
    // trip was exposed on every half row (the epilogue ran at ~2k cycles per half row against 1.2k of MMA work per row).
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;                          // rows alternate between the two epilogue groups
    const int m = quarter * 32 + lane;                          // TMEM lane == MMA row == pixel
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr int kMyRows = kSlots / 2;                         // a group drains every other slot
    int my_slot[kMyRows], my_px[kMyRows];                       // slot (uniform; -1: none) and flat pixel (-1: masked lane) of this warp's rows
#pragma unroll
    for (int k = 0; k < kMyRows; ++k) { my_slot[k] = -1; my_px[k] = -1; }
    for (int q = group; q < sh.order.n_out; q += 2) {           // rows in the order they complete, alternating between the groups
      const int slot = sh.order.out_slot[q];
      int b = 0;
      while (b + 1 < nband && slot >= sh.band[b + 1].slot0) ++b;
      const int j = slot - sh.band[b].slot0;
      const int px0 = sh.lane_px[b][m], pitch = sh.lane_pitch[b][m], nmine = sh.lane_rows[b][m];
#pragma unroll
      for (int k = 0; k < kMyRows; ++k)
        if (k == (q >> 1)) { my_slot[k] = slot; my_px[k] = (px0 >= 0 && j < nmine) ? px0 + j * pitch : -1; }
    }
    uint32_t ndrain[2] = {0, 0};                                // passes drained so far from each TMEM half
    float* const my_bias = sh.bias[warp - 2];
    for (int pass = 0; pass < npass; ++pass) {
      const ConvParams* pp = passes + pass;
      const int dbg = NESR_PROF ? __ldg(&pp->debug_flags) : 0;
      const float* res1 = pp->res1;
      const float* res2 = pp->res2;
      float* dst32a = pp->dst32a;
      float* dst32b = pp->dst32b;
      const float s1 = pp->s1, s2 = pp->s2;
      const int c_off0 = pp->c_off, fmt16 = pp->dst16_fmt, lrelu = pp->lrelu;
      const int coff16 = pp->dst16_coff;
      const int half0 = __ldg(&pp->trunk_half) & 1;             // which 32 columns of a row slot this pass drains
      const int nsub = __ldg(&pp->trunk_no_publish) ? 2 : 1;    // (uniform) conv5's first half: drained with the second (the next pass)
      uint16_t* const base16 = reinterpret_cast<uint16_t*>(pp->dst16) + static_cast<size_t>(coff16 >> 6) * pp->dst16_plane_px * 64 + (coff16 & 63);
      __syncwarp();                                             // every lane is done with the previous pass's bias
      for (int k = lane; k < nsub * COUT; k += 32) my_bias[k] = __ldg(pp->bias + k);   // conv5's halves: consecutive channels of one layer
      const int init = res1 ? 0 : __ldg(&pp->trunk_init);       // (uniform) conv3 / conv4: the drained half receives conv5's bias + residuals
      const ConvParams* p5 = pp + 2;                            // the conv5 pass of that half
      if (init) my_bias[COUT + lane] = __ldg(p5->bias + lane);
      __syncwarp();
      const uint32_t tpar0 = ndrain[half0] & 1u, tpar1 = ndrain[half0 ^ 1] & 1u;
      ++ndrain[half0];
      if (nsub == 2) ++ndrain[half0 ^ 1];
      const bool off = (dbg & 1) != 0;                          // 1: no epilogue loads / stores (timing experiments)

      // TMEM -> registers; the slot is handed back (zeroed, or holding conv5's residuals) AFTER the row's activations are stored
      // and signalled: the stores start the long chain (publish -> acquire -> TMA) the next sweep waits for in every neighbour,
      // the drained slot only a wait inside this CTA.
      auto fetch = [&](int slot, int half, uint32_t par, uint32_t (&r)[COUT / 16][16]) {
        mbar_wait(&sh.tfull[half][slot], par);
        tc_fence_after();
        __syncwarp();
        const uint32_t taddr = lane_base + static_cast<uint32_t>(slot * kSlotCols + half * COUT);
#pragma unroll
        for (int c = 0; c < COUT / 16; ++c) tmem_ld16(taddr + c * 16, r[c]);
      };
      auto hand_back_zeroed = [&](int slot, int half) {
        const uint32_t taddr = lane_base + static_cast<uint32_t>(slot * kSlotCols + half * COUT);
#pragma unroll
        for (int c = 0; c < COUT / 16; ++c) tmem_st16_zero(taddr + c * 16);
      };
      auto handed_back = [&](int slot, int half) {
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&sh.tempty[half][slot]);
      };
      auto add_bias = [&](const uint32_t (&r)[COUT / 16][16], int sub, float (&v)[COUT]) {
        const float4* b4 = reinterpret_cast<const float4*>(my_bias + sub * COUT);
#pragma unroll
        for (int k = 0; k < COUT / 4; ++k) {
          const float4 bv = b4[k];
          v[4 * k] = __uint_as_float(r[(4 * k) >> 4][(4 * k) & 15]) + bv.x;
          v[4 * k + 1] = __uint_as_float(r[(4 * k + 1) >> 4][(4 * k + 1) & 15]) + bv.y;
          v[4 * k + 2] = __uint_as_float(r[(4 * k + 2) >> 4][(4 * k + 2) & 15]) + bv.z;
          v[4 * k + 3] = __uint_as_float(r[(4 * k + 3) >> 4][(4 * k + 3) & 15]) + bv.w;
        }
      };
      auto store16 = [&](int P, int sub, const float (&v)[COUT]) {
        uint32_t w[COUT / 2];
        if (fmt16) {
#pragma unroll
          for (int k = 0; k < COUT / 2; ++k) w[k] = pack2(v[2 * k], v[2 * k + 1], 1);
        } else {
#pragma unroll
          for (int k = 0; k < COUT / 2; ++k) w[k] = pack2(v[2 * k], v[2 * k + 1], 0);
        }
        uint16_t* dst = base16 + static_cast<size_t>(P) * 64 + sub * COUT;
        if (dbg & 16) return;                                    // 16: no 16-bit stores (timing experiments)
        stg256(dst, reinterpret_cast<const uint32_t(&)[8]>(w[0]));
        stg256(dst + 16, reinterpret_cast<const uint32_t(&)[8]>(w[8]));
      };
      auto row_stored = [&](int slot) {                         // this warp's 32 pixels of the row are stored
        if (dbg & 65536) fence_proxy_async_all();                // 65536: a proxy fence in every writer too (timing experiments)
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.stored[slot]);
      };
      // blocked fp32 trunk layout: [pixel/32][ch/8][pixel%32][ch%8]; a lane's 32 channels are 4 runs of 8 floats
      auto trunk_off = [&](int P, int sub) {
        return (static_cast<size_t>(P >> 5) * 8 + ((c_off0 + sub * COUT) >> 3)) * 256 + (static_cast<size_t>(P & 31) << 3);
      };

      if (!res1 && !init) {
        // ---- conv1, conv2 ----
#pragma unroll
        for (int k = 0; k < kMyRows; ++k) {
          const int slot = my_slot[k];
          if (slot < 0) continue;
          uint32_t r[COUT / 16][16];
          fetch(slot, half0, tpar0, r);
          tmem_ld_wait();
          if (my_px[k] >= 0 && !off) {
            float v[COUT];
            add_bias(r, 0, v);
            if (lrelu) {
#pragma unroll
              for (int q = 0; q < COUT; ++q) v[q] = fmaxf(v[q], 0.2f * v[q]);     // LeakyReLU(0.2): slope < 1
            }
            store16(my_px[k], 0, v);
          }
          row_stored(slot);
          hand_back_zeroed(slot, half0);
          handed_back(slot, half0);
        }
      } else if (!res1) {
        // ---- conv3, conv4: drained half <- conv5's bias + res1 / s1 [+ res2 / (s1 s2)] for its 32 channels of that half ----
        // conv5 is  v = (acc + bias) s1 + res1 [; v = v s2 + res2]  =  (acc + bias + res1 / s1 [+ res2 / (s1 s2)]) * s1 [s2]: with the
        // bracket's constant part already in the accumulator, conv5's epilogue loads nothing.  The residual rows are this lane's own
        // pixels (written by its conv5 epilogue of the previous block) and are prefetched one row ahead.
        const float* const i1 = p5->res1;
        const float* const i2 = p5->res2;
        const float inv1 = 1.0f / p5->s1, inv2 = i2 ? inv1 / p5->s2 : 0.0f;
        const int icoff = p5->c_off;
        auto res_off = [&](int P) { return (static_cast<size_t>(P >> 5) * 8 + (icoff >> 3)) * 256 + (static_cast<size_t>(P & 31) << 3); };
        float rt[COUT];
        auto load_res1 = [&](int P) {
          const size_t t = res_off(P);
#pragma unroll
          for (int q = 0; q < COUT / 8; ++q) ldg256(i1 + t + q * 256, &rt[q * 8]);
        };
        if (my_slot[0] >= 0 && my_px[0] >= 0 && !off) load_res1(my_px[0]);
#pragma unroll
        for (int k = 0; k < kMyRows; ++k) {
          const int slot = my_slot[k];
          if (slot < 0) continue;
          const int P = my_px[k];
          const bool on = P >= 0 && !off;
          constexpr int kLast = kMyRows - 1;
          const int Pn = my_px[k < kLast ? k + 1 : k];
          const bool next_row_on = k < kLast && my_slot[k < kLast ? k + 1 : k] >= 0 && Pn >= 0 && !off;
          float r2[COUT];
          if (on && i2) {
            const size_t t = res_off(P);
#pragma unroll
            for (int q = 0; q < COUT / 8; ++q) ldg256_stream(i2 + t + q * 256, &r2[q * 8]);
          }
          uint32_t r[COUT / 16][16];
          fetch(slot, half0, tpar0, r);
          tmem_ld_wait();
          if (on) {
            float v[COUT];
            add_bias(r, 0, v);
            if (lrelu) {
#pragma unroll
              for (int q = 0; q < COUT; ++q) v[q] = fmaxf(v[q], 0.2f * v[q]);
            }
            store16(P, 0, v);
          }
          row_stored(slot);
          const uint32_t taddr = lane_base + static_cast<uint32_t>(slot * kSlotCols + half0 * COUT);
#pragma unroll
          for (int c = 0; c < COUT / 16; ++c) {
            uint32_t iv[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int q = c * 16 + e;
              float x = 0.0f;
              if (on) {
                x = fmaf(rt[q], inv1, my_bias[COUT + q]);
                if (i2) x = fmaf(r2[q], inv2, x);
              }
              iv[e] = __float_as_uint(x);
            }
            tmem_st16(taddr + c * 16, iv);
          }
          handed_back(slot, half0);
          if (next_row_on) load_res1(Pn);
        }
      } else {
        // ---- conv5 (one or both halves per row): the accumulator already holds bias and residuals (see conv3 / conv4) ----
        const float sc = res2 ? s1 * s2 : s1;
#pragma unroll
        for (int k = 0; k < kMyRows; ++k) {
          const int slot = my_slot[k];
          if (slot < 0) continue;
          const int P = my_px[k];
          const bool on = P >= 0 && !off;
          // both halves' 16-bit activations first -- what the next block's first sweep waits for, in every neighbour -- then the
          // row's signal; the fp32 residual (read back by this lane only) is written from a second read of the accumulator
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (sub == 1 && nsub == 1) break;
            uint32_t r[COUT / 16][16];
            fetch(slot, half0 ^ sub, sub ? tpar1 : tpar0, r);
            tmem_ld_wait();
            if (on) {
              float v[COUT];
#pragma unroll
              for (int q = 0; q < COUT; ++q) v[q] = __uint_as_float(r[q >> 4][q & 15]) * sc;
              store16(P, sub, v);
            }
          }
          row_stored(slot);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (sub == 1 && nsub == 1) break;
            if (!(dbg & 4096)) {                                 // 4096: no fp32 trunk stores (timing experiments)
              uint32_t r[COUT / 16][16];                         // (the TMEM load is warp-wide: outside the per-lane mask)
              const uint32_t taddr = lane_base + static_cast<uint32_t>(slot * kSlotCols + (half0 ^ sub) * COUT);
#pragma unroll
              for (int c = 0; c < COUT / 16; ++c) tmem_ld16(taddr + c * 16, r[c]);
              tmem_ld_wait();
              if (on) {
                const size_t toff = trunk_off(P, sub);
#pragma unroll
                for (int q = 0; q < COUT / 8; ++q) {
                  float v[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[(q * 8 + e) >> 4][(q * 8 + e) & 15]) * sc;
                  stg256f(dst32a + toff + q * 256, v);
                  if (dst32b) stg256f_stream(dst32b + toff + q * 256, v);
                }
              }
            }
            hand_back_zeroed(slot, half0 ^ sub);
          }
          handed_back(slot, half0);
          if (nsub == 2) { tc_fence_before(); mbar_arrive(&sh.tempty[half0 ^ 1][slot]); }
        }
      }
      if (threadIdx.x == 64) TS(1, pass + nsub - 1);
      pass += nsub - 1;
    }
  } else {
    // =========================== publisher (warp 10) ===========================
    // Advances this CTA's row counter in the order the rows complete (layout.h trunk_order; the two epilogue groups may
    // finish neighbouring rows out of order).  Rows that are already stored when the previous one is seen are folded into
    // one release: the gpu-scope release is a membar and must not become a per-row cost when the epilogue is ahead.
    if (lane == 0) {
      unsigned* my_word = prog + static_cast<size_t>(blockIdx.x) * kProgStride;
      const int nrows = sh.nrows;
      uint32_t par = 0;
      for (int pass = 0; pass < npass; ++pass) {
        if (__ldg(&passes[pass].trunk_no_publish)) continue;
        int q = 0;
        while (q < nrows) {
          mbar_wait(&sh.stored[sh.order.out_slot[q]], par);
          ++q;
          while (q < nrows && mbar_try_wait(&sh.stored[sh.order.out_slot[q]], par)) ++q;
          if ((dbg0 & 131072) && q < nrows) continue;            // 131072: publish whole passes only (timing experiments)
          // A release is a membar (~1-2k cycles): rows that complete while one is under way are published together by the next.
          // generic-proxy stores of the epilogue warps (observed through the mbarriers) -> async-proxy (TMA) reads of whoever
          // acquires the value released below: the one proxy fence of the chain
          fence_proxy_async_all();
          st_release_gpu(my_word, static_cast<unsigned>(q == nrows ? (pass + 1) * kProgUnit : pass * kProgUnit + q));
        }
        par ^= 1;
        TS(3, pass);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#if NESR_PROF
  if ((dbg0 & 1024) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 40)) {
    const long long t0 = sh.ts[4][0];
    for (int q = 0; q < kTrace && q + ts_base[0] < nsweep; ++q)
      printf("[trunk blk %d sweep %d S%d need=%d] producer_at_first_row %lld  mma_issued %lld\n", (int)blockIdx.x, q + ts_base[0], q % kSweepsPerBlock + 1,
             sweeps[q + ts_base[0]].need, sh.ts[4][q] - t0, sh.ts[0][q] - t0);
    for (int q = 0; q < kTrace && q + ts_base[1] < npass; ++q)
      printf("[trunk blk %d pass %d half %d] epi_rows_done %lld  published %lld\n", (int)blockIdx.x, q + ts_base[1], passes[q + ts_base[1]].trunk_half,
             sh.ts[1][q] - t0, sh.ts[3][q] - t0);
  }
#endif
}

}  // namespace

cudaError_t conv3x3_trunk_configure() {
  return cudaFuncSetAttribute(conv3x3_trunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t launch_conv3x3_trunk(const TrunkMaps& maps, const ConvParams* d_passes, int npass, const TrunkSweep* d_sweeps, int nsweep,
                                 unsigned* d_prog, int grid, cudaStream_t stream) {
  if (grid <= 0 || npass <= 0 || nsweep <= 0) return cudaSuccess;
  if (grid > 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaMemsetAsync(d_prog, 0, static_cast<size_t>(grid) * kProgStride * sizeof(unsigned), stream);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;                 // co-residency: CTAs wait for each other's progress words
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv3x3_trunk_kernel, maps, d_passes, npass, d_sweeps, nsweep, d_prog);
}

}  // namespace nesr
