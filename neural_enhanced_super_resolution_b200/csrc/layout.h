// layout.h -- data layout shared by the host engine and every kernel.
//
// FLAT ZERO-PADDED NHWC.  All tiles of a batch live in one flat pixel array per resolution level
// (level 0 = feature grid h x w, level 1 = 2h x 2w, level 2 = 4h x 4w).  Tile t occupies flat pixels
// [base, base + h*pitch) with pitch = w + 1: pixel (y, x) sits at base + y*pitch + x and column
// x == w is a pad column that is zero and never written.  `base` is a multiple of 128 and at least
// pitch+1 zero pixels separate consecutive tiles.  Consequences:
//   * a 3x3 tap (dy, dx) is the flat shift dy*pitch + dx, and conv zero padding at every tile
//     border -- upstream tile_process forwards each tile independently -- is plain data;
//   * an MMA M-block is 128 consecutive flat pixels (it may span rows), so M tiles are always full;
//   * one 2-D TMA tensor map per buffer serves every tile and every tap.
//
// CHANNEL PLANES.  16-bit activation buffers are stored as planes of 64 channels:
// [plane][pixel][64] (the 192-channel dense-block buffer has three).  A conv reads a channel PREFIX
// of the dense block, i.e. whole planes, so every TMA row slab (136 pixels x 128 B) is one contiguous
// 17 KB run of memory -- interleaving the planes ([pixel][192]) made every 128-byte request hop
// 384 bytes and held HBM efficiency near 45 % (profiles/r1_fold_role_timing.txt).
#pragma once
#include <stdint.h>

namespace nesr {

constexpr int kBlockPixels = 128;       // MMA M
constexpr int kChunkChannels = 64;      // one 128-byte swizzle row of 16-bit channels
constexpr int kFeat = 64;               // num_feat
constexpr int kGrow = 32;               // num_grow_ch
constexpr int kDense = kFeat + 4 * kGrow;   // 192 channels of a dense-block buffer

struct LevelGeom {
  int32_t base;    // flat pixel index of (0,0); multiple of 128
  int32_t pitch;   // w + 1
  int32_t h, w;
};

struct TileGeom {
  LevelGeom lv[3];
  int32_t frame;               // source / destination frame of the batch
  int32_t src_y0, src_x0;      // top-left of the padded tile window in the (pre/mod-padded) input, pixels
  // level-2 pixels [crop_y, crop_y+crop_h) x [crop_x, crop_x+crop_w) are pasted at (out_y0, out_x0)
  int32_t crop_y, crop_x, crop_h, crop_w;
  int32_t out_y0, out_x0;
};

// One unit of conv work: 128 flat pixels starting at `px` (multiple of 128) inside tile `tile`.
struct BlockRef {
  int32_t px;
  int32_t tile;
};

// Row-folded conv kernel work units.  A strip is 128 MMA lanes wide.  A full strip is one segment of
// 128 pixels of one tile.  The narrow right-hand remainder columns of the tiles are cut into vertical
// PIECES and the pieces -- of one tile or of several -- are PACKED side by side into combined strips, each
// with its own halo pixels, so ragged tile widths cost neither MMA lanes nor memory traffic: a 10-pixel
// remainder of a 266-row tile becomes 8 pieces of 34 rows in one 34-row strip instead of a 266-row strip.
struct FoldSeg {
  int32_t tile;    // tile the pixels belong to
  int32_t x0;      // first pixel column
  int32_t width;   // pixels (<= 128)
  int32_t lane0;   // MMA lane of pixel x0 (multiple of 8).  Slab rows [lane0, lane0+width+2) hold pixels x0-1 .. x0+width
  int32_t y0;      // tile row of the segment's strip-row 0 (0 unless the segment is a piece of a cut column)
  int32_t h;       // rows of the segment: strip rows >= h belong to nobody and are dropped
};
constexpr int kMaxFoldSegs = 8;
// One band: output rows [r0, r0+rows) of a strip made of segments segs[seg0 .. seg0+nseg).
struct FoldBand {
  int32_t seg0, nseg, r0, rows;
};

struct ConvParams {
  const BlockRef* blocks;
  const TileGeom* tiles;
  int32_t nblk;
  int32_t level;
  // operands
  const void* src;             // [planes][pixels][64] 16-bit, plane = 64-channel chunk (SIMT kernel reads it directly)
  int32_t src_plane_px;        // pixels per plane of src
  int32_t cin;                 // input channels, multiple of 16
  const void* wpack;           // packed weights arena base (16-bit), rows of 64
  int32_t w_row0;              // first arena row of this layer
  int32_t npad;                // padded Cout = MMA N (16 / 32 / 64)
  int32_t fmt;                 // NESR_FMT_* of src and weights
  uint32_t idesc;
  // epilogue:  v = acc + bias; lrelu?; v = v*s1 + res1; v = v*s2 + res2
  const float* bias;           // [npad]
  int32_t cout;
  int32_t lrelu;
  const float* res1;           // fp32 [pixels][64] or null
  const float* res2;
  float s1, s2;
  float* dst32a;               // fp32 [pixels][64] or null
  float* dst32b;
  void* dst16;                 // 16-bit [planes][pixels][64] or null
  int32_t dst16_plane_px;      // pixels per plane of dst16
  int32_t dst16_coff;
  int32_t dst16_fmt;
  int32_t dst16_up;            // 1: nearest x2 -- write the 2x2 replicas into level+1's layout
  // final layer
  uint8_t* out_u8;             // BGR HWC frames or null
  int64_t out_stride;          // bytes per row
  int64_t out_frame_stride;    // bytes per frame
  int32_t out_trunc;           // out_u8: clip(v * 255, 0, 255) TRUNCATED (the reference HEAD's astype(np.uint8), nesr/nesr.py:897-899) instead of
                               // clamp(v, 0, 1) * 255 rounded half-to-even (upstream RealESRGANer.enhance)
  float* out_f32;              // NCHW frames (unclamped) or null
  int32_t out_h, out_w;        // frame dims for out_f32
  // row-folded kernel (conv3x3_fold.cu)
  const FoldBand* bands;       // all bands of the level, grouped by CTA
  const FoldSeg* segs;         // segments referenced by the bands
  const int32_t* cta_band_off; // [grid + 1] first band of each CTA
  int32_t c_off;               // first output channel of this pass inside the 64-channel fp32 buffers
  int32_t fold_stages;         // activation ring depth (host-computed from the shared-memory budget)
  int32_t src_up;              // 1: `src` is the layer BELOW this level and is read nearest-x2 up-sampled (in[y >> 1, x >> 1]): the tensor maps are
                               // the zero-stride "every pixel twice" views of it, slab rows start one pixel early (at the even pixel x0 - 2)
  // persistent trunk kernel (conv3x3_body.cu)
  int32_t src_sel;             // which of the two dense-block buffers (tensor maps) this pass reads
  int32_t sync_passes;         // passes [0, sync_passes) of every CTA must be complete before this pass loads activations
  int32_t debug_flags;         // NESR_B200_DEBUG_FLAGS (timing experiments only): 1 no epilogue stores, 2 no MMA, 4 no TMA, 8 no TMEM re-zero
  // L2-resident trunk kernel (conv3x3_trunk.cu): input chunk c may be loaded once every CTA has completed passes [0, need[c])
  int32_t need[3];
  int32_t l2_pin_chunks;       // row-folded kernels: loads of input chunks [0, l2_pin_chunks) ask L2 to keep them (0: no hints)
  // [grid][kTrunkMaxDeps] CTAs (its own first) that own a pixel of this CTA's input slab rows; padded with its own index
  const int32_t* trunk_deps;
  // [grid][kTrunkMaxSlabRows][kTrunkMaxDeps] for slab row t of the CTA (its bands in order, input rows -1 .. rows of each) and
  // dependency lane d: how many of its output rows (in the order it completes them) CTA trunk_deps[d] must have stored before the slab row of
  // the newest chunk may be loaded; 0: the slab row holds no pixel of that CTA
  const uint8_t* trunk_need;
  int32_t trunk_no_publish;    // trunk kernel: this pass's rows are not published (conv5 first half: implied by the second half's)
  int32_t trunk_half;          // trunk kernel: which 32-column half of a TMEM row slot this pass's accumulator lives in (0 = A, 1 = B)
  // trunk kernel, conv3 / conv4 passes: instead of re-zeroing the half it has drained, the epilogue leaves conv5's bias and
  // fp32 residuals there -- bias + res1 / s1 [+ res2 / (s1 s2)] of pass + 2, the conv5 pass of that half -- so that conv5's own
  // epilogue is a multiply and stores: its residual loads (L2 round trips, two per channel in every third block) were what
  // the next block's first sweep waited for.  The other conv kernels ignore the field and evaluate the residuals as written.
  int32_t trunk_init;
};

// conv3x3_trunk.cu, MMA side: one chunk sweep over the CTA's bands.  A residual dense block is EIGHT sweeps (SURVEY 8a4:
// conv_k reads cat(x, x1 .. x_{k-1})), merged so that every activation plane is streamed as rarely as possible and the
// MMAs are N = 192 wide (99 % of the tensor rate instead of 85 % at N = 96):
//   S1 x    -> conv1 | conv2      S2 x1    -> conv2          S3 x -> conv3 | conv4      S4 x1,x2 -> conv3 | conv4
//   S5 x3   -> conv4              S6 x     -> conv5          S7 x1,x2 -> conv5          S8 x3,x4 -> conv5
// A TMEM row slot is 64 fp32 columns: half A (conv1, conv3, conv5[0:32]) | half B (conv2, conv4, conv5[32:64]).
struct TrunkSweep {
  int32_t src_sel;             // which dense-block buffer (tensor map pair) holds the plane
  int32_t plane;               // plane index inside that buffer's tensor map
  int32_t w_row0;              // first row of the sweep's [3 dx][nb][64] weight block in the merged-weight arena
  int32_t nb;                  // rows of one dx box: 192 (merged pair / conv5) or 160 (half-B single, zero rows for half A)
  int32_t ks;                  // k-steps of 16 channels (4: whole plane, 2: its first 32 channels)
  int32_t need;                // epilogue passes [0, need) must have stored the rows a slab covers (0: none)
  int32_t flags;               // kSweep* bits
  int32_t pad_;
};
constexpr int kSweepWaitA = 1, kSweepWaitB = 2;       // first toucher of half A / B after a drain: wait for the drained slot
constexpr int kSweepCommitA = 4, kSweepCommitB = 8;   // completes half A / B: commit every output row to the epilogue
constexpr int kSweepSingleB = 16;                     // N = 160 sweep into half B (zero weight rows over half A)
constexpr int kSweepsPerBlock = 8;

// Timing-experiment switches (NESR_B200_DEBUG_FLAGS: results are WRONG when one is set) exist only in a -DNESR_PROF=1 build:
// in the shipped library this folds to 0 and every `dbg & bit` branch in the kernels compiles away.
#ifndef NESR_PROF
#define NESR_PROF 0
#endif
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int dbg_flags(const ConvParams& p) { return NESR_PROF ? p.debug_flags : 0; }

// conv3x3_trunk.cu keeps every output row of a CTA in TMEM for a whole pass: 8 row slots of 64 fp32 columns.
constexpr int kTrunkMaxRows = 8;             // 512 TMEM columns / 64 (two 32-channel accumulators, or conv5's 64 channels, per row)
constexpr int kTrunkMaxBands = 4;
constexpr int kTrunkMaxDeps = 32;            // one polling lane per dependency
constexpr int kTrunkMaxSlabRows = kTrunkMaxRows + 2 * kTrunkMaxBands;   // input rows a CTA streams per chunk sweep

// PHASE ORDER of the trunk kernel's rows (round 2).  Swept top to bottom, the first slab row of a dependent sweep is the LAST
// row the CTA above completes, so no sweep could start before its neighbours' previous pass had ended, however finely rows
// were published (profiles/r2_trunk_experiments.txt, 1. and 5.: ~10k cycles per exposed hand-over, 30k behind conv5).
// Instead every CTA streams its input rows -- halo rows included -- sorted by a key of the TILE row they hold: residue
// y mod 8 in the middle-out sequence 3 4 2 5 1 6 0 7.  An output row is complete once rows y-1, y, y+1 have been swept, so
// rows also COMPLETE in that sequence (it is its own completion order), in every CTA of the group at about the same time
// whatever its row offset: what a sweep needs first is what the previous pass finished first, everywhere, and a row is
// asked for 6-8 row times (~8k cycles) after it was completed -- the hand-over is hidden behind the rest of the sweep.
// Two input rows with the same residue are 8 rows apart and never feed the same output row, so the order in which an output
// row receives its three contributions depends on its tile row alone: results do not depend on how rows are dealt to CTAs.
#if defined(__CUDACC__)
#define NESR_HD __host__ __device__
#else
#define NESR_HD
#endif
NESR_HD inline int trunk_phase(int tile_row) {
  return (0x75310246 >> (4 * (tile_row & 7))) & 7;     // residue 0..7 -> position 6 4 2 0 1 3 5 7
}
struct TrunkOrder {
  int32_t n_in, n_out;
  uint8_t in_band[kTrunkMaxSlabRows];        // processing order of the CTA's input rows: band ...
  uint8_t in_row[kTrunkMaxSlabRows];         // ... and slab row of the band (0: halo above, r + 1: own row r, rows + 1: halo below)
  uint8_t out_slot[kTrunkMaxRows];           // completion order of its output rows (TMEM slot = rows of earlier bands + r)
};
// rows[b], y0[b]: output rows of band b and the tile row of its first one.  Returns false if the CTA does not fit.
NESR_HD inline bool trunk_order(const int* rows, const int* y0, int nband, TrunkOrder& o) {
  o.n_in = 0; o.n_out = 0;
  uint8_t key[kTrunkMaxSlabRows];
  int total = 0;
  if (nband > kTrunkMaxBands) return false;
  for (int b = 0; b < nband; ++b) {
    total += rows[b];
    if (total > kTrunkMaxRows || rows[b] < 1) return false;
    for (int i = 0; i <= rows[b] + 1; ++i) {
      const int k = trunk_phase(y0[b] + i - 1 + 8);            // + 8: the halo above row 0 is row -1
      int at = o.n_in++;
      while (at > 0 && key[at - 1] > k) {                      // stable insertion
        key[at] = key[at - 1]; o.in_band[at] = o.in_band[at - 1]; o.in_row[at] = o.in_row[at - 1];
        --at;
      }
      key[at] = static_cast<uint8_t>(k); o.in_band[at] = static_cast<uint8_t>(b); o.in_row[at] = static_cast<uint8_t>(i);
    }
  }
  uint32_t seen[kTrunkMaxBands] = {0, 0, 0, 0};
  for (int t = 0; t < o.n_in; ++t) {
    const int b = o.in_band[t], i = o.in_row[t];
    seen[b] |= 1u << i;
    int slot0 = 0;
    for (int q = 0; q < b; ++q) slot0 += rows[q];
    for (int j = i - 2; j <= i; ++j)                           // output row j is fed by slab rows j, j + 1, j + 2
      if (j >= 0 && j < rows[b] && ((seen[b] >> j) & 7u) == 7u) o.out_slot[o.n_out++] = static_cast<uint8_t>(slot0 + j);
  }
  return true;
}

}  // namespace nesr
