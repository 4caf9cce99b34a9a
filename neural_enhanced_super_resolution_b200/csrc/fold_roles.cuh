// fold_roles.cuh -- the three warp roles of the row-folded tcgen05 conv, shared by the per-layer
// kernel (conv3x3_fold.cu) and the persistent trunk kernel (conv3x3_body.cu).
//
// Why "folded".  Measured on B200 (tools/umma_probe.cu, profiles/r1_umma_probe_rates.log): a
// tcgen05.mma with M=128, K=16 and both operands in shared memory never issues faster than one per
// ~54 cycles, whatever N <= 64 is -- the 4 KB A-operand read is the cost.  A conv with Cout = 32 as a
// plain implicit GEMM (N = 32) is therefore capped at 30 % of the tensor peak, Cout = 64 at 59 %.
// Folding the three VERTICAL taps into N fixes that: for one input row segment (128 pixels, one A
// tile) and one horizontal shift dx,
//
//      [ out(y-1) | out(y) | out(y+1) ]  +=  A(y, dx) * [ W(dy=+1,dx) | W(dy=0,dx) | W(dy=-1,dx) ]^T
//
// is ONE MMA with N = 3*Cout (96 -> 85 %, 192 -> 99 % of peak in the same probe) whose accumulator
// is three consecutive row slots of a TMEM ring.  Each A byte fetched from shared memory now feeds
// 3x the MACs, and each activation row is loaded from L2 exactly once per strip (no tap re-reads).
//
// Work decomposition.  A tile is cut into 128-pixel-wide column strips and the strips into bands
// of consecutive rows (host: engine.cu build_fold_schedule, balanced over the SMs).  A CTA streams
// down its bands: the TMA producer loads one 136-pixel row slab [136 px x 64 ch] per (row, channel
// chunk) into a ring; the MMA thread issues 3 (dx) x ksteps MMAs per slab, the dx shift being a
// 128-byte offset of the A descriptor into the slab (address-based swizzle makes un-aligned views
// legal: profiles/r1_umma_probe_shifted_views.log); the folded weights of the whole layer pass stay
// resident in shared memory.  Output row r is complete once input row r+1 has been issued, so the
// epilogue warps drain rows in order while the MMAs run ahead: tcgen05.ld -> fused epilogue
// (epilogue.cuh) -> tcgen05.st zeros (every MMA accumulates; a slot is handed back zeroed).
// Rows just outside a band are "virtual": they receive partial sums and are dropped.
//
// One-layer kernel only (end of round 2, profiles/r2_edge_readside.txt):
//   * src_up: the layer reads its input nearest-x2 up-sampled (conv_up1 / conv_up2).  The producer loads through a zero-stride
//     "every pixel twice" tensor map of the layer BELOW, from source row y >> 1; slabs start at the even pixel x0 - 2 and the A
//     descriptors one pixel later.  The TMA engine does the replication: no instruction and no shared-memory traffic on the SM.
//   * the MMA role runs in one thread and looks at the next row's barriers between the MMA groups of the current row;
//   * plain layers (bias, LeakyReLU, one 16-bit store) have a straight-line epilogue; the layer's biases sit in shared memory.
#pragma once
#include <stdio.h>

#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace nesr {
namespace fold {

// NESR_PROF build + debug_flags & 32: block 0 prints where each role spent its cycles (timing experiments only)
#if NESR_PROF
#define PROF_DECL long long prof_t = 0, prof_acc[4] = {0, 0, 0, 0}; const bool prof_on = (dbg_flags(p) & 32) && blockIdx.x == 0
#define PROF_BEGIN() do { if (prof_on) prof_t = clock64(); } while (0)
#define PROF_END(k) do { if (prof_on) { const long long t__ = clock64(); prof_acc[k] += t__ - prof_t; prof_t = t__; } } while (0)
#define PROF_PRINT(...) do { if (prof_on) printf(__VA_ARGS__); } while (0)
#define PROF_NOW() clock64()
#else
#define PROF_DECL do {} while (0)
#define PROF_BEGIN() do {} while (0)
#define PROF_END(k) do {} while (0)
#define PROF_PRINT(...) do {} while (0)
#define PROF_NOW() 0LL
#endif

constexpr int kEpiGroups = 2;                      // epilogue warp-groups (4 warps each) alternating over rows
constexpr int kThreads = 64 + 128 * kEpiGroups;    // warp 0 TMA producer, warp 1 MMA issuer, 8 epilogue warps
constexpr int kSlabPx = 136;                       // 1 + 128 + 1 halo pixels, rounded to 8-row groups
constexpr int kSlabBytes = kSlabPx * 128;          // 17408 = 17 * 1024
constexpr int kMaxStages = 8;
constexpr int kMaxSlots = 16;
constexpr int kSmemBudget = 232448;                // 227 KB
constexpr int kBarrierBytes = (1 + 2 * kMaxStages + 2 * kMaxSlots) * 8 + 32 + 256;   // mbarriers, TMEM slot word (+ pad to 16 B), the layer's 64 biases

template <int COUT>
struct FoldCfg {
  static constexpr int kN3 = 3 * COUT;
  static constexpr int kSlots = (512 / COUT) > kMaxSlots ? kMaxSlots : (512 / COUT);   // 16, 16, 8
  static constexpr int kCols = kSlots * COUT;                                         // 256, 512, 512
  static constexpr int kWBoxBytes = kN3 * 128;                                        // one (dx, chunk) weight box
};

// Shared-memory resources of one CTA: resident folded weights, the activation slab ring and the
// mbarriers of the two pipelines (TMA -> MMA: full/empty per ring stage; MMA -> epilogue:
// tfull/tempty per TMEM row slot).
struct Pipe {
  uint8_t* wsm;
  uint8_t* ring;
  uint64_t *wbar, *full, *empty, *tfull, *tempty;
  uint32_t* tmem_slot;
  float* bias;               // [64] biases of the layer pass (conv3x3_fold.cu copies them once per launch: the per-row epilogue reads them here)
  int nstage;
  uint32_t tmem_base;
};

__device__ __forceinline__ Pipe carve_pipe(uint8_t* smem, int wbytes, int nstage) {
  Pipe s;
  s.wsm = smem;
  s.ring = smem + wbytes;
  s.wbar = reinterpret_cast<uint64_t*>(s.ring + nstage * kSlabBytes);
  s.full = s.wbar + 1;
  s.empty = s.full + kMaxStages;
  s.tfull = s.empty + kMaxStages;
  s.tempty = s.tfull + kMaxSlots;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.tempty + kMaxSlots);
  s.bias = reinterpret_cast<float*>(s.tmem_slot + 6);        // 49 mbarriers = 392 bytes, + 24: 16-byte aligned
  s.nstage = nstage;
  s.tmem_base = 0;
  return s;
}

// Barrier init + TMEM allocation + accumulator zeroing (every MMA accumulates).  All threads call it.
template <int COUT>
__device__ __forceinline__ void pipe_setup(Pipe& s, int warp, int lane, bool zero_tmem) {
  if (warp == 0 && lane == 0) {
    mbar_init(s.wbar, 1);
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&s.tfull[i], 1); mbar_init(&s.tempty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(s.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  s.tmem_base = *s.tmem_slot;
  if (warp >= 2 && warp < 6 && zero_tmem) {
    const uint32_t t0 = s.tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = 0; c < FoldCfg<COUT>::kCols; c += 16) tmem_st16_zero(t0 + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

__device__ __forceinline__ void pipe_teardown(const Pipe& s, int warp) {
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(s.tmem_base, 512);
}

struct RingPos {
  int stage = 0;
  uint32_t phase = 0;
};

// ------------------------------ TMA producer: folded weights of one layer pass ------------------
template <int COUT>
__device__ __forceinline__ void load_weights(const ConvParams& p, const CUtensorMap* wmap, const Pipe& s) {
  const int nchunk = (p.cin + kChunkChannels - 1) / kChunkChannels;
  if (elect_one()) {
    if (dbg_flags(p) & 64) {
      mbar_arrive(s.wbar);
    } else {
      mbar_arrive_expect_tx(s.wbar, 3 * nchunk * FoldCfg<COUT>::kWBoxBytes);
      for (int b = 0; b < 3 * nchunk; ++b)
        tma_load_2d(s.wsm + b * FoldCfg<COUT>::kWBoxBytes, wmap, s.wbar, 0, p.w_row0 + b * FoldCfg<COUT>::kN3);
    }
  }
}

// ------------------------------ TMA producer: activation row slabs (warp-wide, one elected lane issues)
// Band geometry is read once per band; inside a band the slab coordinate advances by the row pitch
// in registers.  (Fetching it per slab made the producer latency-bound on its own metadata loads
// whenever the epilogue kept the memory system busy: profiles/r1_fold_role_timing.txt.)
// Every role copies the ConvParams fields it uses into registers first: in the persistent kernel `p`
// lives in shared memory and each asm memory clobber (every barrier op) would force a reload.
__device__ __forceinline__ void producer_bands(const ConvParams& p, const CUtensorMap* amap, const CUtensorMap* amap8,
                                               const Pipe& s, RingPos& rp, int band_begin, int band_end) {
  const int nchunk = (p.cin + kChunkChannels - 1) / kChunkChannels;
  const int plane_px = p.src_plane_px, level = p.level, dbg = dbg_flags(p), up = p.src_up;
  const FoldBand* const bands = p.bands;
  const FoldSeg* const segs = p.segs;
  const TileGeom* const tiles = p.tiles;
  const int nstage = s.nstage;
  // L2 policy (frame-wide passes: the 213 MB dense-block buffer cannot live in the 126 MB L2, but its first plane --
  // x, read by all six passes of a dense block -- can): chunk 0 loads ask to stay (evict_last), chunks 1, 2 stream.
  // debug_flags & 2048 disables the hints (timing experiments).
  const bool hints = nchunk > 0 && !(dbg & 2048) && p.l2_pin_chunks > 0;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const int pin_chunks = p.l2_pin_chunks;
  PROF_DECL;
  [[maybe_unused]] const long long prof_start = PROF_NOW();
  for (int bi = band_begin; bi < band_end; ++bi) {
    const FoldBand band = bands[bi];
    // per segment: flat pixel of (row r0-1, x0-1), row pitch, slab byte offset, number of 8-pixel boxes
    // read-side nearest x2 (up): the source is the level below; slab row i is source row (r0 - 1 + y0 + i) >> 1 and starts at the EVEN
    // up-sampled pixel x0 - 2 (a "pixel twice" box starts on a pair), one pixel before the usual x0 - 1: seg_px = flat source pixel of
    // (row 0, (x0 - 2) / 2), seg_y0 = up-sampled row of slab row 0.  x0 is even on every level above 0 (tile widths are fw << level).
    int seg_px[kMaxFoldSegs], seg_pitch[kMaxFoldSegs], seg_off[kMaxFoldSegs], seg_n8[kMaxFoldSegs], seg_y0[kMaxFoldSegs];
    bool full_strip = false;
    uint32_t row_bytes = 0;
#pragma unroll
    for (int sgi = 0; sgi < kMaxFoldSegs; ++sgi) {
      seg_px[sgi] = seg_pitch[sgi] = seg_off[sgi] = seg_n8[sgi] = seg_y0[sgi] = 0;
      if (sgi < band.nseg) {
        const FoldSeg sg = segs[band.seg0 + sgi];
        const LevelGeom g = tiles[sg.tile].lv[up ? level - 1 : level];
        seg_px[sgi] = up ? g.base + ((sg.x0 - 2) >> 1) : g.base + (band.r0 - 1 + sg.y0) * g.pitch + sg.x0 - 1;
        seg_y0[sgi] = band.r0 - 1 + sg.y0;
        seg_pitch[sgi] = g.pitch;
        seg_off[sgi] = sg.lane0 * 128;
        seg_n8[sgi] = (sg.width + 2 + up + 7) >> 3;
        if (sg.width == kBlockPixels) full_strip = true;      // a 128-pixel segment is always alone
        row_bytes += seg_n8[sgi] * 1024;
      }
    }
    if (full_strip) row_bytes = kSlabBytes;
    for (int i = 0; i < band.rows + 2; ++i) {
      for (int c = 0; c < nchunk; ++c) {
        PROF_BEGIN();
        mbar_wait(&s.empty[rp.stage], rp.phase ^ 1);
        PROF_END(0);
        if (elect_one()) {
          if (dbg & 4) {
            mbar_arrive(&s.full[rp.stage]);
          } else {
            mbar_arrive_expect_tx(&s.full[rp.stage], row_bytes);
            uint8_t* slab = s.ring + rp.stage * kSlabBytes;
            const int plane = c * plane_px;
            if (up) {                                            // one chunk, no L2 hints: 68 (or 4) source pixels become 136 (8) slab rows
              if (full_strip) {
                tma_load_3d(slab, amap, &s.full[rp.stage], 0, 0, seg_px[0] + ((seg_y0[0] + i) >> 1) * seg_pitch[0]);
              } else {
#pragma unroll
                for (int sgi = 0; sgi < kMaxFoldSegs; ++sgi)
                  for (int k = 0; k < seg_n8[sgi]; ++k)
                    tma_load_3d(slab + seg_off[sgi] + k * 1024, amap8, &s.full[rp.stage], 0, 0,
                                seg_px[sgi] + ((seg_y0[sgi] + i) >> 1) * seg_pitch[sgi] + k * 4);
              }
            } else if (hints) {
              const uint64_t pol = c < pin_chunks ? pol_keep : pol_stream;
              if (full_strip) {
                tma_load_2d_hint(slab, amap, &s.full[rp.stage], 0, plane + seg_px[0] + i * seg_pitch[0], pol);
              } else {
#pragma unroll
                for (int sgi = 0; sgi < kMaxFoldSegs; ++sgi)
                  for (int k = 0; k < seg_n8[sgi]; ++k)
                    tma_load_2d_hint(slab + seg_off[sgi] + k * 1024, amap8, &s.full[rp.stage], 0,
                                     plane + seg_px[sgi] + i * seg_pitch[sgi] + k * 8, pol);
              }
            } else if (full_strip) {
              tma_load_2d(slab, amap, &s.full[rp.stage], 0, plane + seg_px[0] + i * seg_pitch[0]);
            } else {
#pragma unroll
              for (int sgi = 0; sgi < kMaxFoldSegs; ++sgi)
                for (int k = 0; k < seg_n8[sgi]; ++k)
                  tma_load_2d(slab + seg_off[sgi] + k * 1024, amap8, &s.full[rp.stage], 0,
                              plane + seg_px[sgi] + i * seg_pitch[sgi] + k * 8);
            }
          }
        }
        if (++rp.stage == nstage) { rp.stage = 0; rp.phase ^= 1; }
      }
    }
  }
  if ((threadIdx.x & 31) == 0)
    PROF_PRINT("[fold cin=%d] producer: total %lld  wait_empty %lld\n", p.cin, PROF_NOW() - prof_start, prof_acc[0]);
}

// ------------------------------ MMA issuer (warp-wide, one elected lane issues) ------------------
// `u` is the running row-slot counter of the TMEM ring; `wphase` the parity of the weight barrier.
//
// The issuing thread must stay LEAN: the tensor pipe's instruction queue is shallow, so every ~100
// cycles the thread spends outside the issue sequence (a barrier poll, an elect + divergent branch +
// reconvergence -- tools/commit_probe.cu measures ~110 cycles per separately elected commit) is a
// bubble in the MMA stream.  One input row is therefore ONE elected region: the barrier waits for all
// of the row's slabs come first (warp-wide), then a single lane issues the row's 3 x k-steps MMAs of
// every chunk, the commits that free the slabs and the commit that completes the output row.
// kSingle: the caller runs the role in ONE thread (the one-layer kernel): no elect / __syncwarp per row, and the barriers of the NEXT input
// row are looked at (mbarrier.test_wait, non-blocking) between the MMA groups of the current one, so that in the steady state the thread goes
// from a row's last commit straight to the next row's first MMA -- the tensor pipe queues only ~2 MMAs (tools/commit_probe.cu), and the
// warp-wide version left it idle for ~0.9k of the 2.4k cycles a row took on the level-2 edge layers (profiles/r2_edge_readside.txt).
template <int COUT, bool kSingle = false>
__device__ __forceinline__ void mma_bands(const ConvParams& p, const Pipe& s, RingPos& rp, uint32_t& u, uint32_t wphase,
                                          int band_begin, int band_end) {
  using Cfg = FoldCfg<COUT>;
  const int cin = p.cin, dbg = dbg_flags(p);
  const FoldBand* const bands = p.bands;
  const int nchunk = (cin + kChunkChannels - 1) / kChunkChannels;
  const int last_ks = (cin - (nchunk - 1) * kChunkChannels) >> 4;   // k-steps of the last chunk (1..4)
  const uint32_t hw = (p.idesc >> 7) & 7u;                 // operand format bits of the layer
  const uint32_t idesc1 = umma_idesc_f16(hw, COUT), idesc2 = umma_idesc_f16(hw, 2 * COUT), idesc3 = umma_idesc_f16(hw, 3 * COUT);
  const uint32_t hi = umma_desc_hi_sw128();
  const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.ring)) + (p.src_up ? 8u : 0u);   // read-side x2: slabs start one pixel (128 B) early
  const uint32_t w_lo0 = umma_desc_lo(smem_u32(s.wsm));
  const uint32_t tmem_base = s.tmem_base;
  const int nstage = s.nstage;
  const uint32_t dx_lo = nchunk * (Cfg::kWBoxBytes >> 4);  // weight boxes are ordered [dx][chunk]
  constexpr uint32_t kSlabLo = kSlabBytes >> 4, kBoxLo = Cfg::kWBoxBytes >> 4;
  mbar_wait(s.wbar, wphase);
  tc_fence_after();
  PROF_DECL;
  [[maybe_unused]] const long long prof_start = PROF_NOW();
  [[maybe_unused]] int prof_rows = 0;
  for (int bi = band_begin; bi < band_end; ++bi) {
    const int rows = bands[bi].rows;
    prof_rows += rows + 2;
    for (int j = 0; j < 2; ++j) {                          // slots of the first two (virtual) output rows
      const uint32_t v = u + j;
      mbar_wait(&s.tempty[v % Cfg::kSlots], ((v / Cfg::kSlots) & 1) ^ 1);
    }
    bool ready = false;                                    // kSingle: this row's barriers were seen complete during the previous row
    for (int i = 0; i < rows + 2; ++i) {
      PROF_BEGIN();
      if (!ready) {                                        // slot of output row i+2 must be drained + zeroed
        const uint32_t v = u + i + 2;
        mbar_wait(&s.tempty[v % Cfg::kSlots], ((v / Cfg::kSlots) & 1) ^ 1);
      }
      PROF_END(0);
      if (!ready) {                                        // every slab of this input row has landed
        int st = rp.stage;
        uint32_t ph = rp.phase;
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(&s.full[st], ph);
          if (++st == nstage) { st = 0; ph ^= 1; }
        }
      }
      PROF_END(1);
      tc_fence_after();
      const uint32_t q = (u + i) % Cfg::kSlots;
      // accumulator = row slots (q, q+1, q+2); at the ring end it splits into two narrower MMAs
      const uint32_t d0 = tmem_base + q * COUT;
      uint32_t id0 = idesc3, id1 = 0, b1 = 0;
      if (q + 2 == Cfg::kSlots) { id0 = idesc2; id1 = idesc1; b1 = (2 * COUT * 128) >> 4; }
      else if (q + 1 == Cfg::kSlots) { id0 = idesc1; id1 = idesc2; b1 = (COUT * 128) >> 4; }
      if (kSingle || elect_one()) {
        int st = rp.stage;
        for (int c = 0; c < nchunk; ++c) {
          const uint32_t a_lo = a_lo0 + st * kSlabLo;
          const uint32_t b_lo = w_lo0 + c * kBoxLo;
          if (!(dbg & 2)) {
            if (c + 1 < nchunk || last_ks == 4) {
#pragma unroll
              for (int dxi = 0; dxi < 3; ++dxi) {
                if (kSingle && dxi == 2 && c + 1 == nchunk) {   // with two MMA groups queued: are the next row's slot and slabs there?
                  ready = false;
                  if (i + 1 < rows + 2) {
                    const uint32_t v = u + i + 3;
                    ready = mbar_test_wait(&s.tempty[v % Cfg::kSlots], ((v / Cfg::kSlots) & 1) ^ 1);
                    int st2 = rp.stage + nchunk;
                    uint32_t ph2 = rp.phase;
                    if (st2 >= nstage) { st2 -= nstage; ph2 ^= 1; }
                    for (int c2 = 0; c2 < nchunk && ready; ++c2) {
                      ready = mbar_test_wait(&s.full[st2], ph2);
                      if (++st2 == nstage) { st2 = 0; ph2 ^= 1; }
                    }
                  }
                }
                umma_f16_ksteps<4>(d0, a_lo + dxi * 8, b_lo + dxi * dx_lo, hi, id0);
                if (id1) umma_f16_ksteps<4>(tmem_base, a_lo + dxi * 8, b_lo + dxi * dx_lo + b1, hi, id1);
              }
            } else {
#pragma unroll
              for (int dxi = 0; dxi < 3; ++dxi) {
                umma_f16_ksteps_rt(last_ks, d0, a_lo + dxi * 8, b_lo + dxi * dx_lo, hi, id0);
                if (id1) umma_f16_ksteps_rt(last_ks, tmem_base, a_lo + dxi * 8, b_lo + dxi * dx_lo + b1, hi, id1);
              }
            }
          }
          umma_commit(&s.empty[st]);                         // slab may be overwritten once these MMAs have read it
          if (++st == nstage) st = 0;
        }
        umma_commit(&s.tfull[q]);                            // output row i has all its contributions
        if (i == rows + 1) {
          umma_commit(&s.tfull[(u + i + 1) % Cfg::kSlots]);
          umma_commit(&s.tfull[(u + i + 2) % Cfg::kSlots]);
        }
      }
      if (!kSingle) __syncwarp();
      PROF_END(2);
      rp.stage += nchunk;
      if (rp.stage >= nstage) { rp.stage -= nstage; rp.phase ^= 1; }
    }
    u += rows + 4;
  }
  if ((threadIdx.x & 31) == 0)
    PROF_PRINT("[fold cin=%d cout=%d] mma: input rows %d total %lld  wait_tempty %lld  wait_full %lld  issue+commit %lld\n", cin, COUT,
               prof_rows, PROF_NOW() - prof_start, prof_acc[0], prof_acc[1], prof_acc[2]);
}

// ------------------------------ epilogue (warps 2..9) ---------------------------------------------
// kTrunk: a residual-dense-block pass (see EpiRegs).
template <int COUT, bool kTrunk>
__device__ __forceinline__ void epilogue_bands(const ConvParams& p, const Pipe& s, uint32_t& u, int warp, int lane,
                                               int band_begin, int band_end) {
  using Cfg = FoldCfg<COUT>;
  EpiRegs e = make_epi_regs<kTrunk>(p);
  if (!kTrunk) e.bias = s.bias;                              // one-layer launch: biases staged in shared memory (four global loads per 16
                                                             // channels and row sat behind the accumulator wait: profiles/r2_edge_readside.txt)
  const FoldBand* const bands = p.bands;
  const FoldSeg* const segs = p.segs;
  const TileGeom* const tiles = p.tiles;
  const int level = e.level, dbg = dbg_flags(p);
  // The plain layer -- bias, LeakyReLU, one 16-bit store at the layer's own resolution: conv_up1, conv_up2, conv_hr, i.e. every
  // 64-channel layer of the two up-sampled levels -- gets a straight-line epilogue.  The general path re-evaluates a dozen
  // warp-uniform conditions per 16 channels (residuals, fp32 copies, up-sampled store, frame output, operand format): measured on
  // conv_up2 it executed 600 instructions per warp and row, 280 of them branches and uniform compares, and the two epilogue groups --
  // not the MMAs, not the memory system -- set the layer's pace (profiles/r2_edge_readside.txt).
  const bool plain = !kTrunk && e.lrelu && !e.res1 && !e.res2 && !e.dst32a && !e.dst32b && e.dst16 && !e.dst16_up && !e.out_u8 &&
                     !e.out_f32 && (e.dst16_coff & 63) + COUT <= 64 && e.cout == COUT && !(NESR_PROF && (dbg & ~32));   // timing switches other than the role timer use the general path
  uint16_t* const plain_dst = reinterpret_cast<uint16_t*>(e.dst16) + static_cast<size_t>(e.dst16_coff >> 6) * e.dst16_plane_px * 64 + (e.dst16_coff & 63);
  const bool plain_fp16 = e.dst16_fmt != 0;
  const int quarter = warp & 3;
  const int group = (warp - 2) >> 2;                         // rows alternate between the epilogue groups
  const int m = quarter * 32 + lane;                         // A row == TMEM lane == pixel x0 + m
  const uint32_t lane_base = s.tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  PROF_DECL;
  [[maybe_unused]] const long long prof_start = PROF_NOW();
  for (int bi = band_begin; bi < band_end; ++bi) {
    const FoldBand band = bands[bi];
    int my_tile = 0, x = 1 << 20, y_off = 0, y_end = 0;      // this lane's pixel column (none: masked lane), piece rows
    for (int sgi = 0; sgi < band.nseg; ++sgi) {
      const FoldSeg sg = segs[band.seg0 + sgi];
      if (m >= sg.lane0 && m < sg.lane0 + sg.width) { my_tile = sg.tile; x = sg.x0 + (m - sg.lane0); y_off = sg.y0; y_end = sg.h; }
    }
    const TileGeom& tg = tiles[my_tile];
    const LevelGeom g = tg.lv[level];
    const bool lane_on = x < g.w && !(dbg & 1);
    for (int j = 0; j < band.rows + 4; ++j) {
      const uint32_t v = u + j;
      if (static_cast<int>(v % kEpiGroups) != group) continue;
      const uint32_t slot = v % Cfg::kSlots;
      const int ys = band.r0 - 2 + j;                          // strip row; the lane's piece holds strip rows [0, y_end)
      const int y = ys + y_off;
      const bool active = j >= 2 && j < band.rows + 2 && lane_on && ys < y_end;
      PixelRef px;
      px.P = g.base + y * g.pitch + x;
      px.y = y; px.x = x; px.valid = true;
      // residual rows are fetched BEFORE waiting for the accumulator: their latency hides behind the MMAs
      constexpr bool kPrefetchR2 = COUT <= 32;                // 64-wide layers never carry a second residual
      float r1[COUT], r2[kPrefetchR2 ? COUT : 1];
      if (active && e.res1) {
#pragma unroll
        for (int c = 0; c < COUT / 16; ++c) load_trunk16(e.res1, px.P, e.c_off + c * 16, &r1[c * 16]);
      }
      if constexpr (kPrefetchR2) {
        if (active && e.res2) {
#pragma unroll
          for (int c = 0; c < COUT / 16; ++c) load_trunk16(e.res2, px.P, e.c_off + c * 16, &r2[c * 16]);
        }
      }
      PROF_BEGIN();
      mbar_wait(&s.tfull[slot], (v / Cfg::kSlots) & 1);
      PROF_END(0);
      tc_fence_after();
      __syncwarp();
      const uint32_t taddr = lane_base + slot * COUT;
      uint32_t r[COUT / 16][16];
#pragma unroll
      for (int c = 0; c < COUT / 16; ++c) tmem_ld16(taddr + c * 16, r[c]);
      tmem_ld_wait();
      if (!(dbg & 8)) {
#pragma unroll
        for (int c = 0; c < COUT / 16; ++c) tmem_st16_zero(taddr + c * 16);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&s.tempty[slot]);
      PROF_END(1);
      if (plain) {
        if (active) {
          uint16_t* const d = plain_dst + static_cast<size_t>(px.P) * 64;
#pragma unroll
          for (int c = 0; c < COUT / 16; ++c) {
            float v[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 b = reinterpret_cast<const float4*>(s.bias)[c * 4 + q4];
              v[4 * q4 + 0] = __uint_as_float(r[c][4 * q4 + 0]) + b.x; v[4 * q4 + 1] = __uint_as_float(r[c][4 * q4 + 1]) + b.y;
              v[4 * q4 + 2] = __uint_as_float(r[c][4 * q4 + 2]) + b.z; v[4 * q4 + 3] = __uint_as_float(r[c][4 * q4 + 3]) + b.w;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.2f * v[k]);      // LeakyReLU(0.2): v >= 0 ? v : 0.2 v, the same bits without a predicate
            uint32_t w[8];
            if (plain_fp16) {
#pragma unroll
              for (int k = 0; k < 8; ++k) w[k] = pack2(v[2 * k], v[2 * k + 1], 1);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) w[k] = pack2(v[2 * k], v[2 * k + 1], 0);
            }
            stg256(d + c * 16, w);
          }
        }
      } else if (active) {
#pragma unroll
        for (int c = 0; c < COUT / 16; ++c) {
          if (c * 16 < e.cout) {
            float vals[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) vals[k] = __uint_as_float(r[c][k]);
            epilogue16(e, tg, px, c * 16, vals, e.res1 ? &r1[c * 16] : nullptr,
                       (kPrefetchR2 && e.res2) ? &r2[kPrefetchR2 ? c * 16 : 0] : nullptr);
          }
        }
      }
      PROF_END(2);
    }
    u += band.rows + 4;
  }
  if (lane == 0 && (warp == 2 || warp == 6))
    PROF_PRINT("[fold cin=%d cout=%d] epilogue warp %d: total %lld  wait_tfull %lld  ld+zero+arrive %lld  math+stores %lld\n", p.cin,
               COUT, warp, PROF_NOW() - prof_start, prof_acc[0], prof_acc[1], prof_acc[2]);
}

}  // namespace fold
}  // namespace nesr
