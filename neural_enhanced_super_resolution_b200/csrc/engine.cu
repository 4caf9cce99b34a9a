// engine.cu -- host side of libnesr_b200.so: handle, weight repack, tile plan, activation arena,
// layer schedule and the extern "C" entry points declared in include/nesr_b200.h.
//
// What it replaces in the reference stack (Python): RRDBNet construction + load_state_dict,
// RealESRGANer.pre_process / tile_process / post_process / enhance (realesrgan 0.3.0, restated in
// oracle/realesrganer.py) and RRDBNet.forward (basicsr 1.4.2, oracle/rrdbnet.py).
//
// Schedule: ALL tiles of a call (and all frames of a batch) advance through the network together,
// one kernel launch per conv layer, so every launch has thousands of 128-pixel M-blocks to spread
// over the 148 SMs.  The dense-block concat is two ping-pong [pixels][192] buffers written at
// channel offsets; the residual trunk is carried in fp32 ([pixels][64]) next to its 16-bit copy.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/nesr_b200.h"
#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

using namespace nesr;

namespace {

thread_local std::string g_create_error;

struct Layer {
  std::string name;
  int cin = 0, cout = 0;      // true channel counts
  int cin16 = 0;              // cin rounded up to 16 (MMA K step)
  int nchunk = 0;             // 64-channel chunks
  int npad = 0;               // MMA N
  int fmt = 0;                // NESR_FMT_*
  int w_row0 = 0;             // first row in the packed (per-tap) arena
  int bias_off = 0;           // floats into the bias arena
  // row-folded kernel: 1 pass, or 2 passes of 32 output channels when the weights would not fit
  int fold_passes = 1;
  int fold_npad = 0;          // Cout per pass (16 / 32 / 64)
  int fold_row0[2] = {0, 0};  // first row of each pass in the folded arena
};

struct LevelPlan {
  int64_t pixels = 0;         // flat pixels incl. guards
  std::vector<BlockRef> blocks;
  BlockRef* d_blocks = nullptr;
  // row-folded kernel schedule: bands grouped by CTA
  std::vector<FoldBand> bands;
  std::vector<FoldSeg> segs;
  std::vector<int32_t> cta_off;
  FoldBand* d_bands = nullptr;
  FoldSeg* d_segs = nullptr;
  int32_t* d_cta_off = nullptr;
  int fold_grid = 0;
  bool row_capped = false;     // level 0: the schedule was built under the trunk kernel's rules (<= 8 rows per CTA, pieces of a packed strip cut at
                               // multiples of 8 tile rows).  The free schedule a group falls back to may also keep every CTA at <= 8 rows, but its
                               // pieces are not phase aligned: it must never run on the trunk kernel.
  std::vector<int32_t> deps;   // level 0 only: [grid][kTrunkMaxDeps] CTAs owning pixels of a CTA's input slab rows (empty: too many)
  int32_t* d_deps = nullptr;
  std::vector<uint8_t> need;   // level 0 only: [grid][kTrunkMaxSlabRows][kTrunkMaxDeps] rows each dependency must have stored per slab row
  uint8_t* d_need = nullptr;
};

struct PlanKey {
  int n_frames = -1, H = 0, W = 0, tile = 0, tile_pad = 0, pre_pad = 0, first = 0, count = 0, whole = 0, packed = 0;
  std::vector<int32_t> list;             // packed tile-major output of an arbitrary tile subset (slot k = list[k]); empty: the range [first, first+count)
  bool operator==(const PlanKey& o) const {
    return list == o.list && n_frames == o.n_frames && H == o.H && W == o.W && tile == o.tile && tile_pad == o.tile_pad &&
           pre_pad == o.pre_pad && first == o.first && count == o.count && whole == o.whole && packed == o.packed;
  }
};

struct Arena {
  uint8_t* base = nullptr;
  size_t bytes = 0;
  size_t zero_bytes = 0;      // prefix holding every buffer a conv reads (pads must stay zero)
  // sub-buffers
  void* x0 = nullptr;         // [P0][64] 16-bit network input (12 channels used)
  void* d[2] = {nullptr, nullptr};   // [3][P0][64] dense-block ping-pong (three 64-channel planes)
  float* trunk = nullptr;     // [P0][64] fp32 residual trunk
  float* rrdb = nullptr;      // [P0][64] fp32 RRDB input
  float* feat = nullptr;      // [P0][64] fp32 conv_first output (long skip)
  void* g1 = nullptr;         // [P0][64] conv_body's output (+ long skip) on the feature grid: conv_up1 reads it up-sampled (src_up)
  void* g2 = nullptr;         // [P1][64]
  void* g4[2] = {nullptr, nullptr};  // [P2][64]
  int64_t P[3] = {0, 0, 0};
  CUtensorMap m_x0, m_d[2], m_g2, m_g4[2];             // box 128 px (per-tap kernel)
  CUtensorMap f_x0, f_d[2], f_g2, f_g4[2];             // box 136 px (row-folded kernel, full strips)
  CUtensorMap e_x0, e_d[2], e_g2, e_g4[2];             // box 8 px (row-folded kernel, packed remainder strips)
  CUtensorMap u_g1[2], u_g2[2];                        // "every pixel twice" views (zero-stride dimension) of g1 / g2: boxes of 68 and 4 source pixels
  CUtensorMap b_d[2][4];                               // boxes 8 / 16 / 32 / 64 px of the dense-block buffers (trunk kernel)
  CUtensorMap hf_d[2], hb_d[2][4];                     // the same as 32-channel (64-byte, SWIZZLE_64B) boxes: 136 px and 8 / 16 / 32 / 64 px
  bool shared_g = false;                               // growth planes shared by both dense buffers (see build_plan)
};

struct Batch {
  std::vector<TileGeom> tiles;
  TileGeom* d_tiles = nullptr;
  LevelPlan lv[3];
  ConvParams* d_body_passes = nullptr;   // persistent trunk kernel: one ConvParams per RDB layer pass
  int n_body_passes = 0;
  TrunkSweep* d_sweeps = nullptr;        // trunk kernel, MMA side: eight merged chunk sweeps per residual dense block
  int n_sweeps = 0;
  bool trunk_fits = false;               // level-0 schedule fits the TMEM-resident trunk kernel (conv3x3_trunk.cu)
  Arena arena;                           // this group's activation buffers: its own slice of the handle's allocation, or all of it
};


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct nesr_b200_handle {
  nesr_b200_config cfg{};
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;                  // device -> host copies of finished tile groups, overlapped with the next group
  std::vector<cudaEvent_t> ev_group;                   // "group g stitched" events
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc0 = nullptr, evc1 = nullptr;
  // forward_nchw runs on the CALLER's stream but shares the arena, progress words and plan with the entries that run on
  // h->stream: ev_own orders the caller's stream behind our plan uploads, ev_ext orders our next h->stream entry behind the forward
  cudaEvent_t ev_own = nullptr, ev_ext = nullptr;
  bool ext_pending = false;
  std::vector<cudaEvent_t> ev_trunk;                   // begin/end pairs around each trunk kernel launch of the last call
  int n_trunk_timed = 0;
  EncodeTiledFn encode = nullptr;
  std::string error;

  std::vector<Layer> layers;
  std::map<std::string, std::vector<float>> staged;   // checkpoint tensors by name
  std::map<std::string, std::vector<int64_t>> expect; // expected shapes
  bool finalized = false;
  uint16_t* d_wpack = nullptr;
  int64_t w_rows = 0;
  float* d_bias = nullptr;
  CUtensorMap m_w[3];                                  // box rows 16 / 32 / 64
  uint16_t* d_wfold = nullptr;                         // folded weights [dx][chunk][dy: +1,0,-1][cout][64]
  int64_t wfold_rows = 0;
  CUtensorMap m_wf[3];                                 // box rows 48 / 96 / 192
  uint16_t* d_wmerge = nullptr;                        // trunk kernel: per dense block, eight sweep blocks [3 dx][192 | 160][64] (layout.h TrunkSweep)
  int64_t wmerge_rows = 0;
  CUtensorMap m_wm[2];                                 // box rows 192 / 160

  PlanKey key;
  std::vector<Batch> batches;
  uint8_t* arena_base = nullptr;                       // one allocation; every tile group owns a slice, or all share it
  size_t arena_bytes = 0;
  bool arena_shared = false;                           // groups share one slice: its zero pads are re-established per group
  size_t arena_limit = (size_t)24 << 30;               // NESR_B200_ARENA_LIMIT_MB: above this, groups share one slice

  uint8_t* d_in = nullptr;  size_t d_in_bytes = 0;
  uint8_t* d_out = nullptr; size_t d_out_bytes = 0;
  uint8_t* d_tmp = nullptr; size_t d_tmp_bytes = 0;

  unsigned* d_gbar = nullptr;                          // arrival counter of the persistent trunk kernel

  uint8_t* d_pre = nullptr; size_t d_pre_bytes = 0;    // pre-process workspace (Lab planes, CLAHE LUTs)
  int32_t* d_nlm_w[2] = {nullptr, nullptr};            // NLM weight tables (L, ab) for nlm_h[]
  int n_nlm_w[2] = {0, 0};
  float nlm_h[2] = {-1.f, -1.f};

  nesr_b200_stats stats{};
  int debug_flags = 0;        // NESR_B200_DEBUG_FLAGS: timing experiments (results are wrong when set)
  int shared_g = 1;           // NESR_B200_SHARED_G: growth planes of the dense-block buffers single-buffered (trunk kernel path; 0: classic ping-pong of all three planes)
  int l2_pin_chunks = 1;      // NESR_B200_L2_PIN: dense-block planes whose loads are tagged evict_last in the trunk passes
  int max_pieces = 0;         // NESR_B200_MAX_PIECES: remainder pieces packed into one strip of a trunk group (0: the planner picks)
};

namespace {

int fail(nesr_b200_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->error = buf; else g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                        \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return fail(h, NESR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Every C-ABI entry makes the handle's device current for the duration of the call and restores the caller's device on
// every return path (an Engine on cuda:1 must not silently move a torch thread that was on cuda:0).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_SCOPE(h)                                                                          \
  DeviceGuard device_guard__((h)->cfg.device);                                                   \
  if (!device_guard__.ok) return fail(h, NESR_E_CUDA, "cudaSetDevice(%d) failed", (h)->cfg.device)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

uint16_t to16(float v, int fmt) {
  if (fmt == NESR_FMT_FP16) {
    __half hh = __float2half_rn(v);
    uint16_t r; memcpy(&r, &hh, 2); return r;
  }
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  uint16_t r; memcpy(&r, &b, 2); return r;
}

uint32_t hw_fmt(int fmt) { return fmt == NESR_FMT_BF16 ? 1u : 0u; }   // tcgen05 encoding: 0 = f16, 1 = bf16

// ----------------------------------------------------------------------------------------------
// layer table
// ----------------------------------------------------------------------------------------------
void add_layer(nesr_b200_handle* h, const std::string& name, int cin, int cout, int fmt) {
  Layer L;
  L.name = name; L.cin = cin; L.cout = cout; L.fmt = fmt;
  L.cin16 = (int)round_up(cin, 16);
  L.nchunk = (L.cin16 + kChunkChannels - 1) / kChunkChannels;
  L.npad = cout <= 16 ? 16 : (cout <= 32 ? 32 : 64);
  h->layers.push_back(L);
  h->expect[name + ".weight"] = {cout, cin, 3, 3};
  h->expect[name + ".bias"] = {cout};
}

void build_layers(nesr_b200_handle* h) {
  const nesr_b200_config& c = h->cfg;
  const int in_ch = c.feat_in_ch > 0 ? c.feat_in_ch : c.num_in_ch * (c.scale == 2 ? 4 : 1);
  add_layer(h, "conv_first", in_ch, c.num_feat, c.edge_format);
  for (int i = 0; i < c.num_block; ++i)
    for (int j = 1; j <= 3; ++j) {
      const std::string p = "body." + std::to_string(i) + ".rdb" + std::to_string(j) + ".conv";
      for (int k = 1; k <= 4; ++k) add_layer(h, p + std::to_string(k), c.num_feat + (k - 1) * c.num_grow_ch, c.num_grow_ch, c.body_format);
      add_layer(h, p + "5", c.num_feat + 4 * c.num_grow_ch, c.num_feat, c.body_format);
    }
  add_layer(h, "conv_body", c.num_feat, c.num_feat, c.edge_format);
  add_layer(h, "conv_up1", c.num_feat, c.num_feat, c.edge_format);
  add_layer(h, "conv_up2", c.num_feat, c.num_feat, c.edge_format);
  add_layer(h, "conv_hr", c.num_feat, c.num_feat, c.edge_format);
  add_layer(h, "conv_last", c.num_feat, c.num_out_ch, c.edge_format);
  int64_t rows = 0;
  int boff = 0;
  for (Layer& L : h->layers) {
    L.w_row0 = (int)rows;
    rows += (int64_t)9 * L.nchunk * L.npad;
    L.bias_off = boff;
    boff += 64;
  }
  h->w_rows = rows;
  int64_t frows = 0;
  for (Layer& L : h->layers) {
    const bool fits = conv3x3_fold_fits(L.cin16, L.npad);
    L.fold_passes = fits ? 1 : 2;
    L.fold_npad = fits ? L.npad : L.npad / 2;
    for (int ps = 0; ps < L.fold_passes; ++ps) {
      L.fold_row0[ps] = (int)frows;
      frows += (int64_t)3 * L.nchunk * 3 * L.fold_npad;
    }
  }
  h->wfold_rows = frows;
}


// Packed weight layout (16-bit): row ((tap*nchunk + chunk)*npad + n) holds the 64 input channels
// [chunk*64, chunk*64+64) of output channel n at tap (ky,kx) -- K-major rows of 128 B, exactly one
// TMA box [npad x 64] per (tap, chunk).  Padding channels are zero.
void pack_layer(const Layer& L, const float* w_oihw, uint16_t* arena) {
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < L.nchunk; ++c)
      for (int n = 0; n < L.npad; ++n) {
        uint16_t* row = arena + ((int64_t)L.w_row0 + ((int64_t)t * L.nchunk + c) * L.npad + n) * 64;
        for (int k = 0; k < 64; ++k) {
          const int ci = c * 64 + k;
          float v = 0.f;
          if (n < L.cout && ci < L.cin) v = w_oihw[((int64_t)n * L.cin + ci) * 9 + t];
          row[k] = to16(v, L.fmt);
        }
      }
}

// Folded layout of one pass (output channels [n_off, n_off + npad)): row
//   ((dx*nchunk + chunk)*3 + g)*npad + n      g = 0,1,2 <-> dy = +1,0,-1 <-> ky = 2,1,0
// holds input channels [chunk*64, +64) of tap (ky, kx = dx).  One TMA box [3*npad x 64] per (dx, chunk)
// is the B operand of the folded MMA: its three column groups feed output rows y-1, y, y+1.
void pack_layer_fold(const Layer& L, int pass, const float* w_oihw, uint16_t* arena) {
  const int npad = L.fold_npad, n_off = pass * npad;
  for (int dx = 0; dx < 3; ++dx)
    for (int c = 0; c < L.nchunk; ++c)
      for (int g = 0; g < 3; ++g)
        for (int n = 0; n < npad; ++n) {
          uint16_t* row = arena + ((int64_t)L.fold_row0[pass] + (((int64_t)dx * L.nchunk + c) * 3 + g) * npad + n) * 64;
          const int ky = 2 - g, co = n_off + n;
          for (int k = 0; k < 64; ++k) {
            const int ci = c * 64 + k;
            float v = 0.f;
            if (co < L.cout && ci < L.cin) v = w_oihw[((int64_t)co * L.cin + ci) * 9 + ky * 3 + dx];
            row[k] = to16(v, L.fmt);
          }
        }
}

// Merged trunk weights (conv3x3_trunk.cu, layout.h TrunkSweep).  One sweep block is [3 dx][nb rows][64 input channels of one
// plane]; a row is  g*64 + n  with g = 0,1,2 <-> dy = +1,0,-1 <-> ky = 2,1,0  (the three output rows an input row feeds) and
// n the accumulator column inside the 64-column TMEM row slot: n < 32 -> layer `la` (half A), n >= 32 -> layer `lb` (half B);
// conv5 (la == lb, 64 outputs) fills both.  la == nullptr: half-B single sweep -- the 32 rows of half A are zero and the box
// is cut to 160 rows (B | 0 | B | 0 | B starting at column 32 of the first slot).
constexpr int kSweepRows192 = 3 * 192, kSweepRows160 = 3 * 160;
constexpr int kMergeRowsPerBlock = 6 * kSweepRows192 + 2 * kSweepRows160;
const int kSweepRowOff[kSweepsPerBlock] = {0, 576, 1056, 1632, 2208, 2688, 3264, 3840};

void pack_sweep(const Layer* la, const Layer* lb, const float* wa, const float* wb, int plane, int fmt, uint16_t* dst) {
  const bool single = la == nullptr;
  const int nb = single ? 160 : 192;
  for (int dx = 0; dx < 3; ++dx)
    for (int row = 0; row < nb; ++row) {
      const int col = single ? row + 32 : row;                  // accumulator column relative to the first slot's column 0
      const int g = col / 64, n = col % 64;
      const bool conv5 = la == lb;
      const Layer* L = n < 32 ? la : lb;
      const float* w = n < 32 ? wa : wb;
      const int co = conv5 ? n : (n & 31);
      uint16_t* out = dst + ((int64_t)dx * nb + row) * 64;
      for (int k = 0; k < 64; ++k) {
        const int ci = plane * 64 + k;
        float v = 0.f;
        if (L && ci < L->cin && co < L->cout) v = w[((int64_t)co * L->cin + ci) * 9 + (2 - g) * 3 + dx];
        out[k] = to16(v, fmt);
      }
    }
}

// half = true: a box of the FIRST or SECOND 32 channels of the 64-channel pixel rows (inner coordinate 0 or 32), written to shared
// memory as 64-byte rows with the 64-byte swizzle -- the half-width slabs of the trunk kernel's single-layer sweeps.
int make_map(nesr_b200_handle* h, CUtensorMap* m, void* base, int64_t channels, int64_t rows, int box_rows, bool half = false) {
  const cuuint64_t gdim[2] = {(cuuint64_t)channels, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)channels * 2};
  const cuuint32_t box[2] = {half ? 32u : 64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, NESR_E_CUDA, "cuTensorMapEncodeTiled failed (%d) ch=%lld rows=%lld", (int)r,
                                     (long long)channels, (long long)rows);
  return NESR_OK;
}

// (64 channels, 2 copies at stride ZERO, rows pixels): a box of `box_px` source pixels lands in shared memory as 2 * box_px rows of 128
// bytes, every pixel twice -- nearest x2 along x done by the TMA engine (tools/tma_dup_probe.cu).
int make_dup_map(nesr_b200_handle* h, CUtensorMap* m, void* base, int64_t rows, int box_px) {
  const cuuint64_t gdim[3] = {64, 2, (cuuint64_t)rows};
  const cuuint64_t gstride[2] = {0, 128};
  const cuuint32_t box[3] = {64u, 2u, (cuuint32_t)box_px};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, NESR_E_CUDA, "cuTensorMapEncodeTiled (pixel-twice view) failed (%d) rows=%lld", (int)r, (long long)rows);
  return NESR_OK;
}

// ----------------------------------------------------------------------------------------------
// tile plan (upstream RealESRGANer.pre_process + tile_process geometry, oracle/realesrganer.py)
// ----------------------------------------------------------------------------------------------
struct Grid {
  int H2 = 0, W2 = 0, tiles_x = 1, tiles_y = 1;
};

Grid tile_grid_dims(int H, int W, int tile, int pre_pad, int scale) {
  Grid g;
  const int mod = scale == 2 ? 2 : 1;
  g.H2 = (int)round_up(H + pre_pad, mod);
  g.W2 = (int)round_up(W + pre_pad, mod);
  if (tile > 0) {
    g.tiles_x = (g.W2 + tile - 1) / tile;
    g.tiles_y = (g.H2 + tile - 1) / tile;
  }
  return g;
}

void free_batches(nesr_b200_handle* h) {
  for (Batch& b : h->batches) {
    if (b.d_tiles) cudaFree(b.d_tiles);
    if (b.d_body_passes) cudaFree(b.d_body_passes);
    if (b.d_sweeps) cudaFree(b.d_sweeps);
    for (LevelPlan& l : b.lv) {
      if (l.d_blocks) cudaFree(l.d_blocks);
      if (l.d_bands) cudaFree(l.d_bands);
      if (l.d_segs) cudaFree(l.d_segs);
      if (l.d_cta_off) cudaFree(l.d_cta_off);
      if (l.d_deps) cudaFree(l.d_deps);
      if (l.d_need) cudaFree(l.d_need);
    }
  }
  h->batches.clear();
  h->key = PlanKey();
}

void layout_level(Batch& b, int level) {
  int64_t cursor = 0;
  int prev_pitch = 0;
  LevelPlan& lp = b.lv[level];
  lp.blocks.clear();
  for (size_t i = 0; i < b.tiles.size(); ++i) {
    LevelGeom& g = b.tiles[i].lv[level];
    cursor += round_up(std::max(prev_pitch, g.pitch) + 1, kBlockPixels);   // zero guard
    g.base = (int32_t)cursor;
    const int64_t n = (int64_t)g.h * g.pitch;
    for (int64_t j = 0; j < n; j += kBlockPixels) lp.blocks.push_back(BlockRef{(int32_t)(cursor + j), (int32_t)i});
    cursor += round_up(n, kBlockPixels);
    prev_pitch = g.pitch;
  }
  cursor += round_up(prev_pitch + 1, kBlockPixels);
  lp.pixels = cursor;
}

// Row-folded kernel schedule.  Every tile is cut into 128-pixel column strips; the strips of the
// whole batch are laid end to end into one sequence of strip-rows and that sequence is cut into one
// contiguous, equally long run per CTA (split into bands where it crosses a strip boundary).  Every
// CTA gets the same number of rows +-1 and at most a few bands, so no SM waits for a straggler
// (a longest-first deal of fixed-size bands left 12 of 148 CTAs with 35 % more work), and the two
// halo rows a band costs are paid as rarely as possible.
// max_rows > 0 (level 0 of a trunk-kernel group): no CTA gets more output rows than its TMEM row slots; returns false (and
// leaves an incomplete schedule) when the group does not fit under that limit.
bool build_fold_schedule(Batch& b, int level, int num_sms, int max_rows = 0, int max_pieces = 3) {
  LevelPlan& lp = b.lv[level];
  struct Strip { int32_t seg0, nseg, h; };
  std::vector<Strip> strips;
  lp.segs.clear();
  // full 128-pixel strips: one segment each
  for (size_t ti = 0; ti < b.tiles.size(); ++ti) {
    const LevelGeom& g = b.tiles[ti].lv[level];
    for (int x0 = 0; x0 + kBlockPixels <= g.w; x0 += kBlockPixels) {
      strips.push_back(Strip{(int32_t)lp.segs.size(), 1, g.h});
      lp.segs.push_back(FoldSeg{(int32_t)ti, x0, kBlockPixels, 0, 0, g.h});
    }
  }
  // right-hand remainder columns: cut into pieces of at most T rows and packed side by side (with their halo
  // pixels) into combined strips; T is chosen to minimise the number of strip rows (+2 halo rows per strip)
  struct Piece { int32_t tile, x0, rem, span, y0, h; };
  std::vector<Piece> cols;
  for (size_t ti = 0; ti < b.tiles.size(); ++ti) {
    const LevelGeom& g = b.tiles[ti].lv[level];
    const int rem = g.w % kBlockPixels;
    // slab lanes of a piece: its pixels plus one halo pixel on either side -- plus one more on the levels whose layers may read their
    // input nearest-x2 up-sampled (src_up: a slab row then starts at the even pixel x0 - 2), rounded to the 8-pixel TMA box
    if (rem) cols.push_back(Piece{(int32_t)ti, g.w - rem, rem, (rem + 2 + (level > 0 ? 1 : 0) + 7) / 8 * 8, 0, g.h});
  }
  if (!cols.empty()) {
    // every piece costs the TMA producer at least one more operation per slab row; with 8 pieces a packed strip's
    // CTAs were producer-bound and paced the whole group (9.3 instead of 6.7 ms), with 11 one-box operations likewise
    const int kMaxPieces = std::max(1, std::min(max_pieces, kMaxFoldSegs));
    struct Packed { std::vector<Piece> pcs; int lanes = 0, h = 0; };
    auto plan = [&](int T, std::vector<Packed>* out) -> int64_t {
      std::vector<Piece> pcs;
      for (const Piece& c : cols) {
        // trunk groups: pieces start at multiples of 8 tile rows, so that every piece of a packed strip row is in the same
        // phase of the trunk kernel's row order (layout.h trunk_order)
        const int n = (c.h + T - 1) / T, hp = max_rows > 0 ? ((c.h + n - 1) / n + 7) / 8 * 8 : (c.h + n - 1) / n;
        for (int y = 0; y < c.h; y += hp) pcs.push_back(Piece{c.tile, c.x0, c.rem, c.span, y, std::min(hp, c.h - y)});
      }
      std::stable_sort(pcs.begin(), pcs.end(), [](const Piece& a, const Piece& o) { return a.h > o.h; });
      std::vector<Packed> packed;
      for (const Piece& pc : pcs) {
        Packed* dst = nullptr;
        for (Packed& pk : packed)
          if ((int)pk.pcs.size() < kMaxPieces && pk.lanes + pc.span <= 136 && pk.lanes + pc.rem <= kBlockPixels) { dst = &pk; break; }
        if (!dst) { packed.emplace_back(); dst = &packed.back(); }
        dst->pcs.push_back(pc); dst->lanes += pc.span; dst->h = std::max(dst->h, pc.h);
      }
      int64_t cost = 0;
      for (const Packed& pk : packed) cost += pk.h + 2;
      if (out) *out = std::move(packed);
      return cost;
    };
    int best_T = 1 << 30;
    int64_t best = plan(best_T, nullptr);                       // no cutting
    std::vector<int> tried;                                     // a T is evaluated once (same order, same winner: many columns share a height)
    for (const Piece& c : cols)
      for (int n = 2; n <= 16; ++n) {
        const int T = (c.h + n - 1) / n;
        if (T < 8) break;
        if (std::find(tried.begin(), tried.end(), T) != tried.end()) continue;
        tried.push_back(T);
        const int64_t cost = plan(T, nullptr);
        if (cost < best) { best = cost; best_T = T; }
      }
    std::vector<Packed> packed;
    plan(best_T, &packed);
    for (const Packed& pk : packed) {
      Strip st{(int32_t)lp.segs.size(), 0, pk.h};
      int lane = 0;
      for (const Piece& pc : pk.pcs) {
        lp.segs.push_back(FoldSeg{pc.tile, pc.x0, pc.rem, lane, pc.y0, pc.h});
        lane += pc.span;
        ++st.nseg;
      }
      strips.push_back(st);
    }
  }
  int64_t total_rows = 0;
  for (const Strip& st : strips) total_rows += st.h;
  // Tiny groups: the edge-layer kernels (one launch per layer) prefer a few CTAs with >= 4 rows each over 148 CTAs with a
  // row each (per-CTA weight loads); the trunk kernel (max_rows > 0) is latency bound per sweep -- a band of n rows costs
  // n + 2 slab rows in every one of its 552 sweeps -- so its rows are spread over all SMs.
  const int kTrunkMinRows = getenv("NESR_B200_MIN_ROWS") ? atoi(getenv("NESR_B200_MIN_ROWS")) : 1;   // (read per plan: tests vary it)
  const int min_rows = max_rows > 0 ? std::max(1, kTrunkMinRows) : 4;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(num_sms, total_rows / min_rows));
  // Contiguous runs of equal COST: a band of n rows costs n + 2 slab rows (its halo), so a CTA whose run crosses a
  // strip boundary gets fewer output rows.  The smallest per-run budget that covers everything is found by bisection.
  const int nrun = grid;
  std::vector<std::vector<FoldBand>> runs(nrun);
  auto deal = [&](int64_t budget, bool emit) -> bool {
    size_t si = 0;
    int r = 0;                                                 // next row of strips[si]
    for (int c = 0; c < nrun; ++c) {
      if (emit) runs[c].clear();
      int64_t left = budget;
      int rows_left = max_rows > 0 ? max_rows : 1 << 30;
      while (si < strips.size() && left >= 3 && rows_left > 0) {
        const int n = (int)std::min<int64_t>(std::min<int64_t>(left - 2, rows_left), strips[si].h - r);
        if (emit) runs[c].push_back(FoldBand{strips[si].seg0, strips[si].nseg, r, n});
        left -= n + 2;
        rows_left -= n;
        r += n;
        if (r == strips[si].h) { ++si; r = 0; }
      }
    }
    return si == strips.size();
  };
  int64_t lo = 3, hi = total_rows + 2 * (int64_t)strips.size() + 3;
  while (lo < hi) {
    const int64_t mid = (lo + hi) / 2;
    if (deal(mid, false)) hi = mid; else lo = mid + 1;
  }
  const bool complete = deal(lo, true);
  lp.bands.clear();
  lp.cta_off.assign(1, 0);
  for (int c = 0; c < grid; ++c) {
    lp.bands.insert(lp.bands.end(), runs[c].begin(), runs[c].end());
    lp.cta_off.push_back((int32_t)lp.bands.size());
  }
  lp.fold_grid = grid;
  return complete;
}

int build_body_passes(nesr_b200_handle* h, Batch& b);

// conv3x3_trunk.cu keeps all output rows of a CTA in TMEM: at most kTrunkMaxRows rows in kTrunkMaxBands bands.
// Row dependencies of the trunk kernel.  CTA c streams, per chunk sweep, the input rows -1 .. rows of each of its bands
// ("slab rows", pixels x0-1 .. x0+width of every segment).  A slab row of the chunk that another layer pass has just
// written may be loaded once every CTA owning one of its pixels has stored that tile row.  Output: lp.deps[c][d] = the
// CTAs (c itself first) that own such pixels, lp.need[c][t][d] = how many of its output rows (in the order it completes
// them) CTA deps[c][d] must have stored before c loads the t-th slab row of its phase order (layout.h trunk_order).  Strip rows of a packed piece beyond the piece's height
// + 1 feed masked lanes only and carry no dependency; rows outside the tile are zero guard rows.
// Phase order of CTA c's rows (layout.h trunk_order): the kernel derives the same from the same bands.
bool cta_order(const LevelPlan& lp, int c, TrunkOrder& o) {
  int rows[kTrunkMaxBands], y0[kTrunkMaxBands];
  const int b0 = lp.cta_off[c], nb = lp.cta_off[c + 1] - b0;
  if (nb > kTrunkMaxBands) return false;
  for (int b = 0; b < nb; ++b) {
    const FoldBand& band = lp.bands[b0 + b];
    rows[b] = band.rows; y0[b] = band.r0;
    for (int sgi = 0; sgi < band.nseg; ++sgi) {
      const FoldSeg& sg = lp.segs[band.seg0 + sgi];
      if (band.r0 < sg.h) { y0[b] = sg.y0 + band.r0; break; }
    }
  }
  return trunk_order(rows, y0, nb, o);
}

void build_trunk_deps(const std::vector<TileGeom>& tiles, LevelPlan& lp) {
  const int grid = lp.fold_grid;
  struct Rect { int tile, x0, x1, y0, y1, cta, q0; };          // inclusive pixel rectangle of one segment of one band; q0 = slot of tile row y0
  std::vector<Rect> rects;
  std::vector<TrunkOrder> order(grid);
  std::vector<std::array<uint8_t, kTrunkMaxRows>> done_at(grid);   // [cta][slot]: how many rows the CTA has completed once this one is
  for (int c = 0; c < grid; ++c) {
    if (!cta_order(lp, c, order[c])) { lp.deps.clear(); lp.need.clear(); return; }
    done_at[c].fill(0);
    for (int q = 0; q < order[c].n_out; ++q) done_at[c][order[c].out_slot[q]] = (uint8_t)(q + 1);
  }
  for (int c = 0; c < grid; ++c) {
    int slot0 = 0;
    for (int b = lp.cta_off[c]; b < lp.cta_off[c + 1]; ++b) {
      const FoldBand& band = lp.bands[b];
      for (int sgi = 0; sgi < band.nseg; ++sgi) {
        const FoldSeg& sg = lp.segs[band.seg0 + sgi];
        if (band.r0 >= sg.h) continue;                            // this piece has no rows in the band
        rects.push_back(Rect{sg.tile, sg.x0, sg.x0 + sg.width - 1, sg.y0 + band.r0, sg.y0 + std::min(band.r0 + band.rows, sg.h) - 1, c, slot0});
      }
      slot0 += band.rows;
    }
  }
  lp.deps.assign((size_t)grid * kTrunkMaxDeps, 0);
  lp.need.assign((size_t)grid * kTrunkMaxSlabRows * kTrunkMaxDeps, 0);
  for (int c = 0; c < grid; ++c) {
    std::vector<int32_t> dep(1, c);
    for (int t = 0; t < order[c].n_in; ++t) {
      {
        const FoldBand& band = lp.bands[lp.cta_off[c] + order[c].in_band[t]];
        const int i = order[c].in_row[t] - 1;                     // input row of the band: -1 .. rows
        for (int sgi = 0; sgi < band.nseg; ++sgi) {
          const FoldSeg& sg = lp.segs[band.seg0 + sgi];
          const int s = band.r0 + i;                              // strip row of this slab row
          if (band.r0 >= sg.h || s > std::min(band.r0 + band.rows, sg.h)) continue;   // feeds no output row of this piece
          const LevelGeom& g = tiles[sg.tile].lv[0];
          const int y = sg.y0 + s;
          if (y < 0 || y >= g.h) continue;                        // zero guard row
          const int xa = std::max(sg.x0 - 1, 0), xb = std::min(sg.x0 + sg.width, g.w - 1);
          for (const Rect& o : rects) {
            if (o.tile != sg.tile || y < o.y0 || y > o.y1 || o.x1 < xa || o.x0 > xb) continue;
            size_t d = std::find(dep.begin(), dep.end(), o.cta) - dep.begin();
            if (d == dep.size()) {
              if ((int)dep.size() == kTrunkMaxDeps) { lp.deps.clear(); lp.need.clear(); return; }
              dep.push_back(o.cta);
            }
            uint8_t& n = lp.need[((size_t)c * kTrunkMaxSlabRows + t) * kTrunkMaxDeps + d];
            n = std::max<uint8_t>(n, done_at[o.cta][o.q0 + (y - o.y0)]);
          }
        }
      }
    }
    for (int k = 0; k < kTrunkMaxDeps; ++k) lp.deps[(size_t)c * kTrunkMaxDeps + k] = k < (int)dep.size() ? dep[k] : c;
  }
}

bool trunk_schedule_fits(const LevelPlan& lp) {
  for (int c = 0; c < lp.fold_grid; ++c) {
    const int b0 = lp.cta_off[c], b1 = lp.cta_off[c + 1];
    if (b1 - b0 > kTrunkMaxBands) return false;
    int rows = 0;
    for (int b = b0; b < b1; ++b) rows += lp.bands[b].rows;
    if (rows > kTrunkMaxRows) return false;
  }
  return true;
}

// Host half of the plan (no CUDA calls): tiles of the frame(s), split into tile groups, every level's flat layout and
// row-folded schedule, and the trunk kernel's halo dependency lists.  Fills h->batches.
int plan_groups_with_cap(nesr_b200_handle* h, const PlanKey& key, int out_h, int out_w, int64_t default_cap, int pieces);

// Default cap of the product path: a tile group is at most 8 output rows (the trunk kernel's TMEM row slots) of 128 pixels on
// each of the SMs, which also keeps its dense-block working set (~1 KB per feature pixel, about half of it live) inside the
// 126 MB L2.  The schedule decides whether a tile still fits; the pixel cap only prunes the candidates.
// The narrow right-hand remainder columns of the tiles are cut into pieces that are packed side by side into strips; more pieces
// per strip mean fewer strip rows (a 1080p frame: 6 groups with 3 pieces, 5 with 4) but more TMA operations per slab row in the
// CTAs that own them.  Measured (profiles/r2_plan_sweep.txt): a launch costs ~0.27 ms + 0.214 ms per slab row of its busiest CTA,
// plus ~0.12 ms for every piece beyond five.  The planner evaluates 3..6 pieces with that model and keeps the cheapest plan.
int plan_groups(nesr_b200_handle* h, const PlanKey& key, int out_h, int out_w) {
  const int64_t cap = (int64_t)kTrunkMaxRows * kBlockPixels * h->num_sms;
  if (h->cfg.conv_impl != 0 || h->max_pieces > 0) return plan_groups_with_cap(h, key, out_h, out_w, cap, h->max_pieces > 0 ? h->max_pieces : 3);
  double best = 1e30;
  std::vector<Batch> keep;
  for (int pieces = 3; pieces <= 6; ++pieces) {
    h->batches.clear();
    const int rc = plan_groups_with_cap(h, key, out_h, out_w, cap, pieces);
    if (rc != NESR_OK) return rc;
    double score = 0;
    for (const Batch& b : h->batches) {
      int busiest = 0;
      for (int c = 0; c < b.lv[0].fold_grid; ++c) {
        int cost = 0;
        for (int q = b.lv[0].cta_off[c]; q < b.lv[0].cta_off[c + 1]; ++q) cost += b.lv[0].bands[q].rows + 2;
        busiest = std::max(busiest, cost);
      }
      score += 0.27 + 0.214 * busiest + (b.trunk_fits ? 0.12 * std::max(0, pieces - 5) : 1.0 * busiest);   // a whole-frame-kernel group is ~5x slower
    }
    if (getenv("NESR_B200_PLAN_DEBUG")) fprintf(stderr, "nesr_b200 plan: %d pieces per strip: %zu groups, model %.2f ms\n", pieces, h->batches.size(), score);
    if (score < best - 1e-9) { best = score; keep = std::move(h->batches); }
  }
  h->batches = std::move(keep);
  return NESR_OK;
}

int plan_groups_with_cap(nesr_b200_handle* h, const PlanKey& key, int out_h, int out_w, int64_t default_cap, int pieces) {
  const int scale = h->cfg.scale;
  const Grid g = tile_grid_dims(key.H, key.W, key.tile, key.pre_pad, scale);
  std::vector<TileGeom> all;
  const int ntile = g.tiles_x * g.tiles_y;
  const int first = key.whole ? 0 : key.first;
  const int count = key.whole ? ntile : key.count;
  if (first < 0 || count < 0 || first + count > ntile)
    return fail(h, NESR_E_INVALID, "tile range [%d,%d) outside the %d-tile grid", first, first + count, ntile);
  std::vector<int32_t> ids(key.list);
  if (ids.empty()) for (int ti = first; ti < first + count; ++ti) ids.push_back(ti);
  for (int32_t ti : ids) if (ti < 0 || ti >= ntile) return fail(h, NESR_E_INVALID, "tile %d outside the %d-tile grid", ti, ntile);
  for (int f = 0; f < key.n_frames; ++f)
    for (size_t slot = 0; slot < ids.size(); ++slot) {
      const int ti = ids[slot];
      const int ty = ti / g.tiles_x, tx = ti % g.tiles_x;
      int y0 = 0, y1 = g.H2, x0 = 0, x1 = g.W2, y0p = 0, y1p = g.H2, x0p = 0, x1p = g.W2;
      if (key.tile > 0) {
        x0 = tx * key.tile; y0 = ty * key.tile;
        x1 = std::min(x0 + key.tile, g.W2); y1 = std::min(y0 + key.tile, g.H2);
        x0p = std::max(x0 - key.tile_pad, 0); x1p = std::min(x1 + key.tile_pad, g.W2);
        y0p = std::max(y0 - key.tile_pad, 0); y1p = std::min(y1 + key.tile_pad, g.H2);
      }
      const int th = y1p - y0p, tw = x1p - x0p;
      if ((th | tw) & 1)
        return fail(h, NESR_E_INVALID,
                    "tile %d has odd extent %dx%d: pixel_unshuffle(2) needs even tiles (use even tile / tile_pad)", ti, tw, th);
      TileGeom t{};
      t.frame = f; t.src_y0 = y0p; t.src_x0 = x0p;
      const int fh = th / 2, fw = tw / 2;
      for (int l = 0; l < 3; ++l) {
        t.lv[l].h = fh << l; t.lv[l].w = fw << l; t.lv[l].pitch = (fw << l) + 1; t.lv[l].base = 0;
      }
      t.crop_y = (y0 - y0p) * scale; t.crop_x = (x0 - x0p) * scale;
      t.out_y0 = y0 * scale; t.out_x0 = x0 * scale;
      t.crop_h = std::max(0, std::min((y1 - y0) * scale, out_h - t.out_y0));
      t.crop_w = std::max(0, std::min((x1 - x0) * scale, out_w - t.out_x0));
      if (key.packed) {                 // tile-major output: tile k of the range is "frame" k of the slot buffer, pasted at its origin
        t.frame = (int32_t)slot; t.out_y0 = 0; t.out_x0 = 0;
      }
      all.push_back(t);
    }
  // Split into batches (tile groups) bounded by feature pixels.  Product path (conv_impl 0): a group's dense-block
  // working set (~512 B per feature pixel) should stay in the 126 MB L2 across all 414 trunk passes, and its
  // level-0 schedule must fit the TMEM-resident trunk kernel; other paths keep the whole frame in one batch.
  const bool l2_groups = h->cfg.conv_impl == 0;
  const int64_t cap = h->cfg.max_batch_pixels > 0 ? h->cfg.max_batch_pixels : (l2_groups ? default_cap : (int64_t)3 << 20);
  // level-0 schedule of a group and whether the TMEM-resident trunk kernel can run it
  auto sched0 = [&](Batch& bb) {                                // row-capped for the trunk kernel; a group that cannot fit keeps the free schedule
    layout_level(bb, 0);
    bb.lv[0].row_capped = l2_groups && build_fold_schedule(bb, 0, h->num_sms, kTrunkMaxRows, pieces);
    if (!bb.lv[0].row_capped) build_fold_schedule(bb, 0, h->num_sms);
  };
  auto fits0 = [&](const Batch& bb) { return bb.lv[0].row_capped && trunk_schedule_fits(bb.lv[0]); };
  auto tile_px = [&](const TileGeom& t) { return (int64_t)t.lv[0].h * t.lv[0].pitch; };
  std::vector<std::vector<int>> groups;                         // tile indices (into `all`) of every group
  if (l2_groups) {
    // First-fit decreasing: tiles are independent, so a group need not be a contiguous range.  The largest tiles open the
    // groups and the small edge tiles of a frame fill them up (1080p, tile 512: four groups of two 68k-pixel tiles + one
    // 8k-pixel bottom-row tile each, instead of a tail group with almost no work per CTA and 414 dependent passes of latency).
    std::vector<int> order(all.size());
    for (size_t i = 0; i < all.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tile_px(all[x]) > tile_px(all[y]); });
    // A probe builds the group's whole level-0 schedule; with hundreds of small tiles (tile 128 on a 4K frame: 510) probing every open
    // group for every tile took minutes.  Strip rows a tile needs AT LEAST -- its full strips, plus its remainder column's lanes at the
    // 136 slab lanes of a packed strip row -- bound a group from below: a group whose bound exceeds the TMEM row budget cannot fit.
    auto min_rows = [&](const TileGeom& t) {
      const int w = t.lv[0].w, h = t.lv[0].h, rem = w % kBlockPixels;
      return (int64_t)h * (w / kBlockPixels) + (rem ? (int64_t)h * ((rem + 2 + 7) / 8 * 8) / 136 : 0);   // rounded DOWN: a bound, never an estimate
    };
    const int64_t row_budget = (int64_t)kTrunkMaxRows * h->num_sms;
    std::vector<int64_t> gpx, grows;
    for (int ti : order) {
      const int64_t n = tile_px(all[ti]), nr = min_rows(all[ti]);
      bool placed = false;
      for (size_t g = 0; g < groups.size() && !placed; ++g) {
        if (gpx[g] + n > cap) continue;
        if (h->cfg.max_batch_pixels == 0 && grows[g] + nr > row_budget) continue;      // (an explicit pixel cap keeps the probes: groups may
                                                                                       // exceed the trunk kernel's budget there on purpose)
        Batch probe;
        for (int k : groups[g]) probe.tiles.push_back(all[k]);
        probe.tiles.push_back(all[ti]);
        sched0(probe);
        if (!fits0(probe)) continue;
        groups[g].push_back(ti); gpx[g] += n; grows[g] += nr; placed = true;
      }
      if (!placed) { groups.push_back({ti}); gpx.push_back(n); grows.push_back(nr); }
    }
    for (auto& g : groups) std::sort(g.begin(), g.end());
    std::sort(groups.begin(), groups.end(), [](const std::vector<int>& x, const std::vector<int>& y) { return x.front() < y.front(); });
  } else {
    int64_t px = 0;
    for (size_t i = 0; i < all.size(); ++i) {
      const int64_t n = tile_px(all[i]);
      if (groups.empty() || px + n > cap) { groups.emplace_back(); px = 0; }
      groups.back().push_back((int)i);
      px += n;
    }
  }
  for (const auto& g : groups) {
    Batch b;
    for (int k : g) b.tiles.push_back(all[k]);
    for (int l = 1; l < 3; ++l) { layout_level(b, l); build_fold_schedule(b, l, h->num_sms); }
    sched0(b);
    b.trunk_fits = l2_groups && fits0(b);
    if (b.trunk_fits) {
      build_trunk_deps(b.tiles, b.lv[0]);
      b.trunk_fits = !b.lv[0].deps.empty();
    }
    if (b.lv[2].pixels >= ((int64_t)1 << 31) - 4096) return fail(h, NESR_E_INVALID, "batch too large for 32-bit pixel indices");
    if (getenv("NESR_B200_PLAN_DEBUG")) {
      int64_t px = 0, rows = 0;
      for (const TileGeom& t : b.tiles) px += tile_px(t);
      for (const FoldBand& fb : b.lv[0].bands) rows += fb.rows;
      int max_cost = 0;
      for (int c = 0; c < b.lv[0].fold_grid; ++c) {
        int cost = 0;
        for (int q = b.lv[0].cta_off[c]; q < b.lv[0].cta_off[c + 1]; ++q) cost += b.lv[0].bands[q].rows + 2;
        max_cost = std::max(max_cost, cost);
      }
      fprintf(stderr, "nesr_b200 plan: group %zu: %zu tiles, %lld px, %lld strip rows on %d CTAs, %zu bands, max slab rows per CTA %d, trunk %d (deps %s)\n", h->batches.size(),
              b.tiles.size(), (long long)px, (long long)rows, b.lv[0].fold_grid, b.lv[0].bands.size(), max_cost, (int)b.trunk_fits, b.lv[0].deps.empty() ? "none" : "ok");
    }
    h->batches.push_back(std::move(b));
  }
  return NESR_OK;
}

int build_plan(nesr_b200_handle* h, const PlanKey& key, int out_h, int out_w) {
  if (h->key == key && !h->batches.empty()) return NESR_OK;
  free_batches(h);
  if (int prc = plan_groups(h, key, out_h, out_w)) return prc;
  // upload
  int64_t P[3] = {0, 0, 0};
  for (Batch& b : h->batches) {
    CUDA_TRY(h, cudaMalloc(&b.d_tiles, b.tiles.size() * sizeof(TileGeom)));
    CUDA_TRY(h, cudaMemcpyAsync(b.d_tiles, b.tiles.data(), b.tiles.size() * sizeof(TileGeom), cudaMemcpyHostToDevice, h->stream));
    for (int l = 0; l < 3; ++l) {
      LevelPlan& lp = b.lv[l];
      CUDA_TRY(h, cudaMalloc(&lp.d_blocks, std::max<size_t>(1, lp.blocks.size()) * sizeof(BlockRef)));
      CUDA_TRY(h, cudaMemcpyAsync(lp.d_blocks, lp.blocks.data(), lp.blocks.size() * sizeof(BlockRef), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(h, cudaMalloc(&lp.d_bands, std::max<size_t>(1, lp.bands.size()) * sizeof(FoldBand)));
      CUDA_TRY(h, cudaMemcpyAsync(lp.d_bands, lp.bands.data(), lp.bands.size() * sizeof(FoldBand), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(h, cudaMalloc(&lp.d_segs, std::max<size_t>(1, lp.segs.size()) * sizeof(FoldSeg)));
      CUDA_TRY(h, cudaMemcpyAsync(lp.d_segs, lp.segs.data(), lp.segs.size() * sizeof(FoldSeg), cudaMemcpyHostToDevice, h->stream));
      if (!lp.need.empty()) {
        CUDA_TRY(h, cudaMalloc(&lp.d_need, lp.need.size()));
        CUDA_TRY(h, cudaMemcpyAsync(lp.d_need, lp.need.data(), lp.need.size(), cudaMemcpyHostToDevice, h->stream));
      }
      if (!lp.deps.empty()) {
        CUDA_TRY(h, cudaMalloc(&lp.d_deps, lp.deps.size() * sizeof(int32_t)));
        CUDA_TRY(h, cudaMemcpyAsync(lp.d_deps, lp.deps.data(), lp.deps.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
      }
      CUDA_TRY(h, cudaMalloc(&lp.d_cta_off, lp.cta_off.size() * sizeof(int32_t)));
      CUDA_TRY(h, cudaMemcpyAsync(lp.d_cta_off, lp.cta_off.data(), lp.cta_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
      P[l] = std::max(P[l], lp.pixels);
    }
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));      // host vectors may now be reused
  // Activation arena.  Every tile group gets its own slice (its zero pads -- pad columns, guard rows -- are written once and
  // never touched again); only when that would need more than 24 GB do all groups share one slice sized for the largest,
  // whose pads are then re-established by a memset before every group (0.2 ms per group at 1080p).
  auto slice_bytes = [&](const Batch& bb, size_t* zero) {
    const int64_t p0 = bb.lv[0].pixels, p1 = bb.lv[1].pixels, p2 = bb.lv[2].pixels;
    const size_t sz_x0 = round_up(p0 * 64 * 2, 1024), sz_d = round_up(p0 * kDense * 2, 1024);
    const size_t sz_f = round_up(p0 * 64 * 4, 1024), sz_g2 = round_up(p1 * 64 * 2, 1024), sz_g4 = round_up(p2 * 64 * 2, 1024);
    if (zero) *zero = 2 * sz_x0 + 2 * sz_d + sz_g2 + 2 * sz_g4;
    return 2 * sz_x0 + 2 * sz_d + 3 * sz_f + sz_g2 + 2 * sz_g4;            // x0 and g1 have the same size
  };
  size_t total = 0, largest = 0;
  for (const Batch& bb : h->batches) { const size_t n = slice_bytes(bb, nullptr); total += n; largest = std::max(largest, n); }
  h->arena_shared = h->batches.size() > 1 && total > h->arena_limit;
  // a shared slice must hold the largest extent of EVERY level (groups differ in shape)
  size_t need = total;
  if (h->arena_shared) {
    Batch big;
    for (int l = 0; l < 3; ++l) big.lv[l].pixels = P[l];
    need = slice_bytes(big, nullptr);
  }
  if (need > h->arena_bytes) {
    if (h->arena_base) cudaFree(h->arena_base);
    h->arena_base = nullptr; h->arena_bytes = 0;
    cudaError_t e = cudaMalloc(&h->arena_base, need);
    if (e != cudaSuccess) return fail(h, NESR_E_NOMEM, "activation arena of %zu bytes: %s", need, cudaGetErrorString(e));
    h->arena_bytes = need;
  }
  int rc;
  uint8_t* cursor = h->arena_base;
  for (Batch& bb : h->batches) {
    Arena& a = bb.arena;
    int64_t Pb[3];
    for (int l = 0; l < 3; ++l) Pb[l] = h->arena_shared ? P[l] : bb.lv[l].pixels;
    const size_t sz_x0 = round_up(Pb[0] * 64 * 2, 1024), sz_d = round_up(Pb[0] * kDense * 2, 1024);
    const size_t sz_f = round_up(Pb[0] * 64 * 4, 1024), sz_g2 = round_up(Pb[1] * 64 * 2, 1024), sz_g4 = round_up(Pb[2] * 64 * 2, 1024);
    uint8_t* p = h->arena_shared ? h->arena_base : cursor;
    a.base = p;
    a.x0 = p; p += sz_x0;
    a.g1 = p; p += sz_x0;
    // Dense-block buffers.  Classic layout: two [3][P0][64] buffers that ping-pong between blocks.  Shared-growth layout (trunk
    // kernel only, NESR_B200_SHARED_G): only x needs the ping-pong -- a block's conv1 overwrites the previous block's growth
    // channels only after every halo neighbour has published that block's conv5, i.e. has finished reading them -- so the planes
    // are [xB][xA][x1|x2][x3|x4]: buffer A = planes 1..3 (contiguous, as before), buffer B = planes 0, 2, 3 (the kernel skips one
    // plane after chunk 0).  20 % fewer live bytes per pixel in L2.
    a.shared_g = h->shared_g && h->cfg.conv_impl == 0 && bb.trunk_fits;
    if (a.shared_g) {
      const size_t plane = (size_t)Pb[0] * 64 * 2;
      a.d[1] = p; a.d[0] = p + plane; p += round_up((int64_t)(4 * plane), 1024);
    } else {
      a.d[0] = p; p += sz_d; a.d[1] = p; p += sz_d;
    }
    a.g2 = p; p += sz_g2;
    a.g4[0] = p; p += sz_g4; a.g4[1] = p; p += sz_g4;
    a.zero_bytes = (size_t)(p - a.base);              // every buffer a conv reads: pads must be zero
    a.trunk = (float*)p; p += sz_f; a.rrdb = (float*)p; p += sz_f; a.feat = (float*)p; p += sz_f;
    a.bytes = (size_t)(p - a.base);
    cursor = p;
    for (int l = 0; l < 3; ++l) a.P[l] = Pb[l];
    if (!h->arena_shared || &bb == &h->batches.front()) CUDA_TRY(h, cudaMemsetAsync(a.base, 0, a.zero_bytes, h->stream));
    if ((rc = make_map(h, &a.m_x0, a.x0, 64, Pb[0], kBlockPixels))) return rc;
    const int64_t dpx[2] = {3 * Pb[0], (a.shared_g ? 4 : 3) * Pb[0]};     // buffer B spans four planes in the shared-growth layout
    for (int i2 = 0; i2 < 2; ++i2) if ((rc = make_map(h, &a.m_d[i2], a.d[i2], 64, dpx[i2], kBlockPixels))) return rc;
    if ((rc = make_map(h, &a.m_g2, a.g2, 64, Pb[1], kBlockPixels))) return rc;
    for (int i2 = 0; i2 < 2; ++i2) if ((rc = make_map(h, &a.m_g4[i2], a.g4[i2], 64, Pb[2], kBlockPixels))) return rc;
    constexpr int kSlab = 136;                                 // row slab of the folded kernel
    if ((rc = make_map(h, &a.f_x0, a.x0, 64, Pb[0], kSlab))) return rc;
    for (int i2 = 0; i2 < 2; ++i2) if ((rc = make_map(h, &a.f_d[i2], a.d[i2], 64, dpx[i2], kSlab))) return rc;
    if ((rc = make_map(h, &a.f_g2, a.g2, 64, Pb[1], kSlab))) return rc;
    for (int i2 = 0; i2 < 2; ++i2) if ((rc = make_map(h, &a.f_g4[i2], a.g4[i2], 64, Pb[2], kSlab))) return rc;
    if ((rc = make_map(h, &a.e_x0, a.x0, 64, Pb[0], 8))) return rc;
    for (int i2 = 0; i2 < 2; ++i2) if ((rc = make_map(h, &a.e_d[i2], a.d[i2], 64, dpx[i2], 8))) return rc;
    for (int i2 = 0; i2 < 2; ++i2)
      for (int k = 0; k < 4; ++k) if ((rc = make_map(h, &a.b_d[i2][k], a.d[i2], 64, dpx[i2], 8 << k))) return rc;
    for (int i2 = 0; i2 < 2; ++i2) {
      if ((rc = make_map(h, &a.hf_d[i2], a.d[i2], 64, dpx[i2], kSlab, true))) return rc;
      for (int k = 0; k < 4; ++k) if ((rc = make_map(h, &a.hb_d[i2][k], a.d[i2], 64, dpx[i2], 8 << k, true))) return rc;
    }
    if ((rc = make_map(h, &a.e_g2, a.g2, 64, Pb[1], 8))) return rc;
    for (int i2 = 0; i2 < 2; ++i2) if ((rc = make_map(h, &a.e_g4[i2], a.g4[i2], 64, Pb[2], 8))) return rc;
    if ((rc = make_dup_map(h, &a.u_g1[0], a.g1, Pb[0], kSlab / 2)) || (rc = make_dup_map(h, &a.u_g1[1], a.g1, Pb[0], 4))) return rc;
    if ((rc = make_dup_map(h, &a.u_g2[0], a.g2, Pb[1], kSlab / 2)) || (rc = make_dup_map(h, &a.u_g2[1], a.g2, Pb[1], 4))) return rc;
  }
  h->stats.arena_bytes = (int64_t)h->arena_bytes;
  if (h->cfg.conv_impl == 0 || h->cfg.conv_impl == 4)
    for (Batch& b : h->batches)
      if ((rc = build_body_passes(h, b))) return rc;
  h->key = key;
  return NESR_OK;
}

// ----------------------------------------------------------------------------------------------
// network schedule
// ----------------------------------------------------------------------------------------------
struct ConvIO {
  const CUtensorMap* amap = nullptr;     // box 128 px
  const CUtensorMap* fmap = nullptr;     // box 136 px
  const CUtensorMap* emap = nullptr;     // box 8 px
  const void* src = nullptr;
  int planes = 1;
  int level = 0;
};

// Fills the layer-dependent fields of `p` (operands, weights, bias) for layer L read through `io`.
void bind_layer(nesr_b200_handle* h, const Batch& b, const Layer& L, const ConvIO& io, ConvParams& p) {
  const LevelPlan& lp = b.lv[io.level];
  p.blocks = lp.d_blocks; p.tiles = b.d_tiles; p.nblk = (int)lp.blocks.size(); p.level = io.level;
  p.src = io.src; p.src_plane_px = (int)b.arena.P[io.level]; p.cin = L.cin16;
  p.wpack = h->d_wpack; p.w_row0 = L.w_row0; p.npad = L.npad; p.fmt = L.fmt;
  p.idesc = umma_idesc_f16(hw_fmt(L.fmt), (uint32_t)L.npad);
  p.bias = h->d_bias + L.bias_off; p.cout = L.cout;
  p.bands = lp.d_bands; p.segs = lp.d_segs; p.cta_band_off = lp.d_cta_off;
  p.debug_flags = h->debug_flags;
}

// Narrows a bound ConvParams to pass `ps` of the row-folded kernel (Cout split when the folded weights
// of the whole layer would not fit in shared memory).
void select_fold_pass(const Layer& L, int ps, const ConvParams& bound, ConvParams& p) {
  p = bound;
  p.npad = L.fold_npad;
  p.c_off = ps * L.fold_npad;
  p.bias = bound.bias + p.c_off;
  p.dst16_coff = bound.dst16_coff + p.c_off;
  p.cout = std::min(L.cout - p.c_off, L.fold_npad);
  p.w_row0 = L.fold_row0[ps];
  p.idesc = umma_idesc_f16((bound.idesc >> 7) & 7u, (uint32_t)L.fold_npad);
}

const CUtensorMap& fold_weight_map(const nesr_b200_handle* h, int fold_npad) {
  return h->m_wf[fold_npad == 16 ? 0 : (fold_npad == 32 ? 1 : 2)];
}

int run_conv(nesr_b200_handle* h, const Batch& b, const Layer& L, const ConvIO& io, ConvParams p, cudaStream_t s) {
  const LevelPlan& lp = b.lv[io.level];
  bind_layer(h, b, L, io, p);
  cudaError_t e = cudaSuccess;
  if (h->cfg.conv_impl == 1) {
    e = launch_conv3x3_simt(p, s);
  } else {
    for (int ps = 0; ps < L.fold_passes && e == cudaSuccess; ++ps) {
      ConvParams q;
      select_fold_pass(L, ps, p, q);
      e = launch_conv3x3_fold(*io.fmap, *io.emap, fold_weight_map(h, L.fold_npad), q, lp.fold_grid, s);
      h->stats.conv_launches++;
      if (ps + 1 < L.fold_passes) h->stats.kernel_launches++;
    }
  }
  h->stats.kernel_launches++;
  if (e != cudaSuccess) return fail(h, NESR_E_CUDA, "conv %s launch failed: %s", L.name.c_str(), cudaGetErrorString(e));
  return NESR_OK;
}

// Epilogue wiring of the 5 convs of residual dense block r (0-based over the whole trunk) reading
// dense buffer `cur`: conv1-4 append 32 channels to the same buffer (torch.cat as addressing), conv5
// writes 0.2*x5 + x (+ the RRDB skip on every third block) as the next block's channels [0,64).
ConvParams rdb_conv_params(const nesr_b200_handle* h, const Arena& a, int r, int k, int cur) {
  const nesr_b200_config& c = h->cfg;
  const int nrdb = c.num_block * 3;
  ConvParams p{};
  p.dst16_plane_px = (int)a.P[0];
  if (k <= 4) {                      // x_k = lrelu(conv_k(cat(x, x1..x_{k-1})))  -> channels [64+32(k-1), +32)
    p.lrelu = 1;
    p.dst16 = a.d[a.shared_g ? 0 : cur]; p.dst16_coff = kFeat + (k - 1) * kGrow; p.dst16_fmt = c.body_format;   // shared growth planes sit behind xA
  } else {                           // x5*0.2 + x  (+ RRDB skip on every third block)
    // fp32 residuals: Y = a.trunk carries the running sum inside an RRDB, X = a.rrdb holds the RRDB's input (the first RRDB's is
    // conv_first's output, a.feat, which the long skip needs again).  The block that closes an RRDB reads Y and the RRDB input and
    // writes its result ONCE, into X (a lane reads its own channels before it overwrites them): that value is both the next RRDB's
    // input and its first block's residual, which therefore reads X and writes Y.  (Round 1 wrote it twice, into trunk and rrdb.)
    const int i = r / 3, j = r % 3;
    p.s1 = 0.2f;
    p.res1 = (j == 0 && i > 0) ? a.rrdb : a.trunk;
    p.dst32a = a.trunk;
    if (j == 2) {
      p.res2 = (i == 0) ? a.feat : a.rrdb; p.s2 = 0.2f;
      p.dst32a = a.rrdb;
    }
    p.dst16 = a.d[cur ^ 1]; p.dst16_coff = 0;
    p.dst16_fmt = (r == nrdb - 1) ? c.edge_format : c.body_format;    // conv_body reads the last one
  }
  return p;
}

// One ConvParams per layer pass of the trunk, in execution order, for the persistent kernel.
int build_body_passes(nesr_b200_handle* h, Batch& b) {
  const Arena& a = b.arena;
  const int nrdb = h->cfg.num_block * 3;
  std::vector<ConvParams> passes;
  std::vector<TrunkSweep> sweeps;
  size_t li = 1;                     // layers[0] is conv_first
  int cur = 0;
  for (int r = 0; r < nrdb; ++r) {
    const ConvIO io{&a.m_d[cur], &a.f_d[cur], &a.e_d[cur], a.d[cur], 3, 0};
    for (int k = 1; k <= 5; ++k) {
      const Layer& L = h->layers[li++];
      if (L.fold_npad != 32 || L.fmt != h->cfg.body_format)
        return fail(h, NESR_E_STATE, "persistent trunk kernel: layer %s is not a 32-channel pass", L.name.c_str());
      ConvParams bound = rdb_conv_params(h, b.arena, r, k, cur);
      bind_layer(h, b, L, io, bound);
      const int first = (int)passes.size();
      for (int ps = 0; ps < L.fold_passes; ++ps) {
        ConvParams q;
        select_fold_pass(L, ps, bound, q);
        q.src_sel = cur | (a.shared_g ? 2 : 0);      // bit 1: buffer B's chunks 1, 2 are one plane further (shared-growth layout)
        q.sync_passes = first;       // passes of one layer read the same input and write disjoint channels
        // trunk kernel: which passes must be complete everywhere before input chunk c may be loaded.  Chunk 0 is x
        // (previous block's conv5, both halves), chunk 1 holds x1|x2 (conv1, conv2), chunk 2 holds x3|x4 (conv3, conv4).
        const int rb = r * 6;
        q.need[0] = rb;
        q.need[1] = k == 2 ? rb + 1 : rb + 2;
        q.need[2] = k == 4 ? rb + 3 : rb + 4;
        q.trunk_deps = b.lv[0].d_deps;
        q.trunk_need = b.lv[0].d_need;
        // the first half of conv5 is needed by nobody before the second half has been published too: skip its publish
        q.trunk_no_publish = (k == 5 && ps + 1 < L.fold_passes) ? 1 : 0;
        q.trunk_half = k == 5 ? ps : ((k - 1) & 1);      // TMEM half: conv1, conv3, conv5[0:32] -> A; conv2, conv4, conv5[32:64] -> B
        q.trunk_init = (k == 3 || k == 4) ? 1 : 0;       // its drained half receives conv5's bias + residuals (pass + 2)
        q.l2_pin_chunks = h->l2_pin_chunks;
        passes.push_back(q);
      }
    }
    {  // MMA side of the trunk kernel: the block's eight merged sweeps.  `need` counts epilogue passes (six per block).
      const int pb = r * 6, wb = r * kMergeRowsPerBlock;
      auto plane = [&](int c) { return c + ((a.shared_g && cur == 1 && c > 0) ? 1 : 0); };   // buffer B: xB | (xA) | x1x2 | x3x4
      const TrunkSweep blk[kSweepsPerBlock] = {
          {cur, plane(0), wb + kSweepRowOff[0], 192, 4, pb, kSweepWaitA | kSweepWaitB | kSweepCommitA, 0},
          {cur, plane(1), wb + kSweepRowOff[1], 160, 2, pb + 1, kSweepWaitA | kSweepCommitB | kSweepSingleB, 0},
          {cur, plane(0), wb + kSweepRowOff[2], 192, 4, pb, kSweepWaitB, 0},
          {cur, plane(1), wb + kSweepRowOff[3], 192, 4, pb + 2, kSweepCommitA, 0},
          {cur, plane(2), wb + kSweepRowOff[4], 160, 2, pb + 3, kSweepWaitA | kSweepCommitB | kSweepSingleB, 0},
          {cur, plane(0), wb + kSweepRowOff[5], 192, 4, pb, kSweepWaitB, 0},
          {cur, plane(1), wb + kSweepRowOff[6], 192, 4, pb + 2, 0, 0},
          {cur, plane(2), wb + kSweepRowOff[7], 192, 4, pb + 4, kSweepCommitA | kSweepCommitB, 0},
      };
      sweeps.insert(sweeps.end(), blk, blk + kSweepsPerBlock);
    }
    cur ^= 1;
  }
  if (b.d_body_passes) cudaFree(b.d_body_passes);
  b.d_body_passes = nullptr;
  b.n_body_passes = (int)passes.size();
  CUDA_TRY(h, cudaMalloc(&b.d_body_passes, passes.size() * sizeof(ConvParams)));
  CUDA_TRY(h, cudaMemcpyAsync(b.d_body_passes, passes.data(), passes.size() * sizeof(ConvParams), cudaMemcpyHostToDevice, h->stream));
  if (b.d_sweeps) cudaFree(b.d_sweeps);
  b.d_sweeps = nullptr;
  b.n_sweeps = (int)sweeps.size();
  CUDA_TRY(h, cudaMalloc(&b.d_sweeps, sweeps.size() * sizeof(TrunkSweep)));
  CUDA_TRY(h, cudaMemcpyAsync(b.d_sweeps, sweeps.data(), sweeps.size() * sizeof(TrunkSweep), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return NESR_OK;
}

struct Sink {
  uint8_t* out_u8 = nullptr; int64_t out_stride = 0, out_frame_stride = 0;
  int out_trunc = 0;
  float* out_f32 = nullptr; int out_h = 0, out_w = 0;
};

int forward_batch(nesr_b200_handle* h, const Batch& b, const PackParams& pack_in, const Sink& sink, cudaStream_t s,
                  bool time_begin, bool time_end, bool time_trunk = false) {
  const Arena& a = b.arena;
  const nesr_b200_config& c = h->cfg;
  int rc;
  if (h->arena_shared) {         // groups share one slice but have different flat layouts: re-establish the zero pads
    cudaError_t ez = cudaMemsetAsync(a.base, 0, a.zero_bytes, s);
    if (ez != cudaSuccess) return fail(h, NESR_E_CUDA, "arena memset failed: %s", cudaGetErrorString(ez));
  }
  PackParams pk = pack_in;
  pk.blocks = b.lv[0].d_blocks; pk.tiles = b.d_tiles; pk.nblk = (int)b.lv[0].blocks.size();
  pk.x0 = a.x0; pk.fmt = c.edge_format;
  cudaError_t e = launch_pack(pk, s);
  h->stats.kernel_launches++;
  if (e != cudaSuccess) return fail(h, NESR_E_CUDA, "pack launch failed: %s", cudaGetErrorString(e));
  if (time_begin) cudaEventRecord(h->evc0, s);

  size_t li = 0;
  auto next = [&]() -> const Layer& { return h->layers[li++]; };
  {  // conv_first: x0 -> trunk (fp32), feat (fp32), d[0][0:64] (16-bit copy for the first RDB)
    ConvParams p{};
    p.dst32a = a.trunk; p.dst32b = a.feat;
    p.dst16 = a.d[0]; p.dst16_plane_px = (int)a.P[0]; p.dst16_coff = 0; p.dst16_fmt = c.body_format;
    if ((rc = run_conv(h, b, next(), ConvIO{&a.m_x0, &a.f_x0, &a.e_x0, a.x0, 1, 0}, p, s))) return rc;
  }
  int cur = 0;
  const int nrdb = c.num_block * 3;
  if (c.conv_impl == 0 || c.conv_impl == 4) {   // all RDB layer passes in one persistent cooperative launch
    TrunkMaps tm;
    if (b.trunk_fits) {
      for (int i2 = 0; i2 < 2; ++i2) {
        tm.full[i2] = a.f_d[i2];
        for (int k = 0; k < 4; ++k) { tm.box[i2][k] = a.b_d[i2][k]; tm.hbox[i2][k] = a.hb_d[i2][k]; }
        tm.half[i2] = a.hf_d[i2];
      }
      tm.w192 = h->m_wm[0]; tm.w160 = h->m_wm[1];
    }
    if (time_trunk) {                  // events on the launching stream around the dominant kernel
      while ((int)h->ev_trunk.size() < 2 * (h->n_trunk_timed + 1)) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreate(&ev) != cudaSuccess) return fail(h, NESR_E_CUDA, "cudaEventCreate failed");
        h->ev_trunk.push_back(ev);
      }
      cudaEventRecord(h->ev_trunk[2 * h->n_trunk_timed], s);
    }
    cudaError_t eb = b.trunk_fits
        ? launch_conv3x3_trunk(tm, b.d_body_passes, b.n_body_passes, b.d_sweeps, b.n_sweeps, h->d_gbar, b.lv[0].fold_grid, s)
        : launch_conv3x3_body(a.f_d[0], a.f_d[1], a.e_d[0], a.e_d[1], fold_weight_map(h, 32), b.d_body_passes,
                              b.n_body_passes, h->d_gbar, b.lv[0].fold_grid, s);
    if (eb != cudaSuccess) return fail(h, NESR_E_CUDA, "trunk kernel launch failed: %s", cudaGetErrorString(eb));
    if (time_trunk) { cudaEventRecord(h->ev_trunk[2 * h->n_trunk_timed + 1], s); ++h->n_trunk_timed; }
    h->stats.kernel_launches++;
    h->stats.conv_launches++;
    li += (size_t)nrdb * 5;
    cur = nrdb & 1;
  } else {
    for (int r = 0; r < nrdb; ++r) {
      const ConvIO io{&a.m_d[cur], &a.f_d[cur], &a.e_d[cur], a.d[cur], 3, 0};
      for (int k = 1; k <= 5; ++k)
        if ((rc = run_conv(h, b, next(), io, rdb_conv_params(h, b.arena, r, k, cur), s))) return rc;
      cur ^= 1;
    }
  }
  // nearest x2 (upstream F.interpolate(scale_factor=2, mode='nearest') in front of conv_up1 and conv_up2) on the READ side: the producing
  // layer stores its own resolution and the consuming layer's TMA loads deliver every source pixel twice (zero-stride tensor-map
  // dimension) from source row y >> 1 -- a quarter of the stores and loads of the write-side form (four replicated stores), which the
  // CUDA-core validation kernel (conv_impl 1: it reads `src` directly) keeps.  Index arithmetic only: same bits either way.
  const bool read_up = c.conv_impl != 1;
  const void* last_src = nullptr;
  const CUtensorMap *last_m = nullptr, *last_f = nullptr, *last_e = nullptr;
  if (read_up) {
    {  // conv_body + long skip -> g1 (feature grid)
      ConvParams p{};
      p.res1 = a.feat; p.s1 = 1.0f;
      p.dst16 = a.g1; p.dst16_plane_px = (int)a.P[0]; p.dst16_fmt = c.edge_format;
      if ((rc = run_conv(h, b, next(), ConvIO{&a.m_d[cur], &a.f_d[cur], &a.e_d[cur], a.d[cur], 3, 0}, p, s))) return rc;
    }
    {  // conv_up1 + lrelu on level 1, reading g1 up-sampled -> g2
      ConvParams p{};
      p.lrelu = 1; p.src_up = 1;
      p.dst16 = a.g2; p.dst16_plane_px = (int)a.P[1]; p.dst16_fmt = c.edge_format;
      if ((rc = run_conv(h, b, next(), ConvIO{nullptr, &a.u_g1[0], &a.u_g1[1], a.g1, 1, 1}, p, s))) return rc;
    }
    {  // conv_up2 + lrelu on level 2, reading g2 up-sampled -> g4[0]
      ConvParams p{};
      p.lrelu = 1; p.src_up = 1;
      p.dst16 = a.g4[0]; p.dst16_plane_px = (int)a.P[2]; p.dst16_fmt = c.edge_format;
      if ((rc = run_conv(h, b, next(), ConvIO{nullptr, &a.u_g2[0], &a.u_g2[1], a.g2, 1, 2}, p, s))) return rc;
    }
    {  // conv_hr + lrelu: g4[0] -> g4[1]
      ConvParams p{};
      p.lrelu = 1;
      p.dst16 = a.g4[1]; p.dst16_plane_px = (int)a.P[2]; p.dst16_fmt = c.edge_format;
      if ((rc = run_conv(h, b, next(), ConvIO{&a.m_g4[0], &a.f_g4[0], &a.e_g4[0], a.g4[0], 1, 2}, p, s))) return rc;
    }
    last_src = a.g4[1]; last_m = &a.m_g4[1]; last_f = &a.f_g4[1]; last_e = &a.e_g4[1];
  } else {
    {  // conv_body + long skip, stored nearest-x2 upsampled into level 1
      ConvParams p{};
      p.res1 = a.feat; p.s1 = 1.0f;
      p.dst16 = a.g2; p.dst16_plane_px = (int)a.P[1]; p.dst16_fmt = c.edge_format; p.dst16_up = 1;
      if ((rc = run_conv(h, b, next(), ConvIO{&a.m_d[cur], &a.f_d[cur], &a.e_d[cur], a.d[cur], 3, 0}, p, s))) return rc;
    }
    {  // conv_up1 + lrelu, stored upsampled into level 2
      ConvParams p{};
      p.lrelu = 1;
      p.dst16 = a.g4[0]; p.dst16_plane_px = (int)a.P[2]; p.dst16_fmt = c.edge_format; p.dst16_up = 1;
      if ((rc = run_conv(h, b, next(), ConvIO{&a.m_g2, &a.f_g2, &a.e_g2, a.g2, 1, 1}, p, s))) return rc;
    }
    {  // conv_up2 + lrelu
      ConvParams p{};
      p.lrelu = 1;
      p.dst16 = a.g4[1]; p.dst16_plane_px = (int)a.P[2]; p.dst16_fmt = c.edge_format;
      if ((rc = run_conv(h, b, next(), ConvIO{&a.m_g4[0], &a.f_g4[0], &a.e_g4[0], a.g4[0], 1, 2}, p, s))) return rc;
    }
    {  // conv_hr + lrelu
      ConvParams p{};
      p.lrelu = 1;
      p.dst16 = a.g4[0]; p.dst16_plane_px = (int)a.P[2]; p.dst16_fmt = c.edge_format;
      if ((rc = run_conv(h, b, next(), ConvIO{&a.m_g4[1], &a.f_g4[1], &a.e_g4[1], a.g4[1], 1, 2}, p, s))) return rc;
    }
    last_src = a.g4[0]; last_m = &a.m_g4[0]; last_f = &a.f_g4[0]; last_e = &a.e_g4[0];
  }
  {  // conv_last -> clamp, BGR, u8, halo crop + stitch  (or unclamped fp32 NCHW)
    ConvParams p{};
    p.out_u8 = sink.out_u8; p.out_stride = sink.out_stride; p.out_frame_stride = sink.out_frame_stride; p.out_trunc = sink.out_trunc;
    p.out_f32 = sink.out_f32; p.out_h = sink.out_h; p.out_w = sink.out_w;
    if ((rc = run_conv(h, b, next(), ConvIO{last_m, last_f, last_e, last_src, 1, 2}, p, s))) return rc;
  }
  if (time_end) cudaEventRecord(h->evc1, s);
  h->stats.tiles_processed += (int64_t)b.tiles.size();
  return NESR_OK;
}

// Called at the start of every entry that works on h->stream: a forward_nchw still running on the caller's stream uses the
// same arena and progress words.
int wait_external(nesr_b200_handle* h) {
  if (!h->ext_pending) return NESR_OK;
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_ext, 0));
  h->ext_pending = false;
  return NESR_OK;
}

int ensure(nesr_b200_handle* h, uint8_t** buf, size_t* cap, size_t need) {
  if (need <= *cap) return NESR_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc(buf, need);
  if (e != cudaSuccess) return fail(h, NESR_E_NOMEM, "device image buffer of %zu bytes: %s", need, cudaGetErrorString(e));
  *cap = need;
  return NESR_OK;
}

int enhance_impl(nesr_b200_handle* h, const uint8_t* in, int n_frames, int H, int W, int64_t in_stride,
                 int64_t in_frame_stride, int tile, int tile_pad, int pre_pad, int first, int count, int whole,
                 uint8_t* out, int64_t out_stride, int64_t out_frame_stride, int flags, int packed = 0,
                 const int32_t* tile_ids = nullptr, int head = 0) {
  if (!h) return NESR_E_INVALID;
  if (!h->finalized) return fail(h, NESR_E_STATE, "weights not finalized");
  if (h->cfg.feat_in_ch > 0 && h->cfg.feat_in_ch != 12)
    return fail(h, NESR_E_STATE, "this handle's conv_first takes %d feature channels (a scale-4 / scale-1 network): images go through nesr_b200_forward_feat_f32", h->cfg.feat_in_ch);
  if (!in || !out || n_frames < 1 || H < 2 || W < 2) return fail(h, NESR_E_INVALID, "bad image arguments (H=%d W=%d n=%d)", H, W, n_frames);
  if (tile < 0 || tile_pad < 0 || pre_pad < 0 || pre_pad >= H || pre_pad >= W)
    return fail(h, NESR_E_INVALID, "bad tile/pad arguments (tile=%d tile_pad=%d pre_pad=%d)", tile, tile_pad, pre_pad);
  if (in_stride < (int64_t)(head ? W / 2 : W) * 3 || (!packed && out_stride < (int64_t)W * h->cfg.scale * 3)) return fail(h, NESR_E_INVALID, "row stride smaller than a row");
  if (packed && (n_frames != 1 || !(flags & NESR_PTR_OUT_DEVICE))) return fail(h, NESR_E_INVALID, "tile-major output needs one frame and a device buffer");
  DEVICE_SCOPE(h);
  if (int wrc = wait_external(h)) return wrc;
  const int s = h->cfg.scale, OH = H * s, OW = W * s;
  PlanKey key; key.n_frames = n_frames; key.H = H; key.W = W; key.tile = tile; key.tile_pad = tile_pad;
  key.pre_pad = pre_pad; key.first = first; key.count = count; key.whole = whole; key.packed = packed;
  if (tile_ids) { key.list.assign(tile_ids, tile_ids + count); key.first = 0; key.count = 0; }
  int rc = build_plan(h, key, OH, OW);
  if (rc) return rc;

  const uint8_t* d_in = in;
  int64_t d_in_stride = in_stride, d_in_fs = in_frame_stride;
  const int in_h = head ? H / 2 : H, in_w = head ? W / 2 : W;      // HEAD mode: (H, W) is the virtual frame whose un-shuffle the 12 channels are
  if (!(flags & NESR_PTR_IN_DEVICE)) {
    d_in_stride = (int64_t)in_w * 3; d_in_fs = d_in_stride * in_h;
    if ((rc = ensure(h, &h->d_in, &h->d_in_bytes, (size_t)d_in_fs * n_frames))) return rc;
    for (int f = 0; f < n_frames; ++f)
      CUDA_TRY(h, cudaMemcpy2DAsync(h->d_in + f * d_in_fs, d_in_stride, in + f * in_frame_stride, in_stride, (size_t)in_w * 3, in_h,
                                    cudaMemcpyHostToDevice, h->stream));
    d_in = h->d_in;
  }
  uint8_t* d_out = out;
  int64_t d_out_stride = out_stride, d_out_fs = out_frame_stride;
  if (!(flags & NESR_PTR_OUT_DEVICE)) {
    d_out_stride = (int64_t)OW * 3; d_out_fs = d_out_stride * OH;
    if ((rc = ensure(h, &h->d_out, &h->d_out_bytes, (size_t)d_out_fs * n_frames))) return rc;
    d_out = h->d_out;
    if (!whole)   // partial tile range: untouched pixels must round-trip unchanged
      for (int f = 0; f < n_frames; ++f)
        CUDA_TRY(h, cudaMemcpy2DAsync(d_out + f * d_out_fs, d_out_stride, out + f * out_frame_stride, out_stride, (size_t)OW * 3, OH,
                                      cudaMemcpyHostToDevice, h->stream));
  }
  cudaEventRecord(h->ev0, h->stream);
  h->n_trunk_timed = 0;
  PackParams pk{};
  pk.in_u8 = d_in; pk.in_stride = d_in_stride; pk.in_frame_stride = packed ? 0 : d_in_fs; pk.H = H; pk.W = W; pk.pre_pad = pre_pad;
  if (head) { pk.in_u8 = nullptr; pk.in_u8_head = d_in; pk.head_replicate = head == 2; }
  Sink sink;
  sink.out_trunc = head ? 1 : 0; sink.out_u8 = d_out; sink.out_stride = d_out_stride; sink.out_frame_stride = d_out_fs;
  // Host output of a whole frame: every tile group's stitched rectangles go back on a second stream while the next group
  // computes (the tiles' crop rectangles partition the frame).  A tile range keeps the single full-frame copy.
  const bool overlap_d2h = !(flags & NESR_PTR_OUT_DEVICE) && whole && h->batches.size() > 1 && h->copy_stream;
  for (size_t bi = 0; bi < h->batches.size(); ++bi) {
    if ((rc = forward_batch(h, h->batches[bi], pk, sink, h->stream, bi == 0, bi + 1 == h->batches.size(), true))) return rc;
    if (overlap_d2h) {
      while (h->ev_group.size() <= bi) {
        cudaEvent_t ev = nullptr;
        CUDA_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->ev_group.push_back(ev);
      }
      CUDA_TRY(h, cudaEventRecord(h->ev_group[bi], h->stream));
      CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_group[bi], 0));
      for (const TileGeom& t : h->batches[bi].tiles) {
        if (t.crop_w <= 0 || t.crop_h <= 0) continue;
        const size_t off_d = (size_t)t.frame * d_out_fs + (size_t)t.out_y0 * d_out_stride + (size_t)t.out_x0 * 3;
        const size_t off_h = (size_t)t.frame * out_frame_stride + (size_t)t.out_y0 * out_stride + (size_t)t.out_x0 * 3;
        CUDA_TRY(h, cudaMemcpy2DAsync(out + off_h, out_stride, d_out + off_d, d_out_stride, (size_t)t.crop_w * 3, t.crop_h,
                                      cudaMemcpyDeviceToHost, h->copy_stream));
      }
    }
  }
  cudaEventRecord(h->ev1, h->stream);
  if (overlap_d2h) {
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
  } else if (!(flags & NESR_PTR_OUT_DEVICE)) {
    for (int f = 0; f < n_frames; ++f)
      CUDA_TRY(h, cudaMemcpy2DAsync(out + f * out_frame_stride, out_stride, d_out + f * d_out_fs, d_out_stride, (size_t)OW * 3, OH,
                                    cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.last_device_ms = ms;
  if (cudaEventElapsedTime(&ms, h->evc0, h->evc1) == cudaSuccess) h->stats.last_conv_ms = ms;
  double trunk_ms = 0;
  for (int t = 0; t < h->n_trunk_timed; ++t)
    if (cudaEventElapsedTime(&ms, h->ev_trunk[2 * t], h->ev_trunk[2 * t + 1]) == cudaSuccess) trunk_ms += ms;
  h->stats.last_trunk_ms = trunk_ms;
  h->stats.last_trunk_launches = h->n_trunk_timed;
  return NESR_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

void nesr_b200_default_config(nesr_b200_config* cfg, int32_t device) {
  if (!cfg) return;
  memset(cfg, 0, sizeof *cfg);
  cfg->abi_version = NESR_B200_ABI_VERSION;
  cfg->device = device;
  cfg->num_in_ch = 3; cfg->num_out_ch = 3; cfg->scale = 2;
  cfg->num_feat = 64; cfg->num_block = 23; cfg->num_grow_ch = 32;
  cfg->body_format = NESR_FMT_BF16; cfg->edge_format = NESR_FMT_FP16;
  cfg->conv_impl = 0; cfg->max_batch_pixels = 0;
}

int nesr_b200_create(const nesr_b200_config* cfg, nesr_b200_handle** out) {
  if (!cfg || !out) return fail(nullptr, NESR_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != NESR_B200_ABI_VERSION) return fail(nullptr, NESR_E_INVALID, "ABI version %d != %d", cfg->abi_version, NESR_B200_ABI_VERSION);
  if (cfg->scale != 2 || cfg->num_feat != 64 || cfg->num_grow_ch != 32 || cfg->num_in_ch != 3 || cfg->num_out_ch != 3 || cfg->num_block < 1)
    return fail(nullptr, NESR_E_INVALID, "unsupported architecture: this build implements RRDBNet(3,3,scale=2,num_feat=64,num_grow_ch=32)");
  if (cfg->feat_in_ch < 0 || cfg->feat_in_ch > 64) return fail(nullptr, NESR_E_INVALID, "feat_in_ch %d outside 0..64", cfg->feat_in_ch);
  if ((cfg->body_format | cfg->edge_format) & ~1) return fail(nullptr, NESR_E_INVALID, "bad operand format");
  if (cfg->conv_impl != 0 && cfg->conv_impl != 1 && cfg->conv_impl != 3 && cfg->conv_impl != 4)
    return fail(nullptr, NESR_E_INVALID, "conv_impl %d: 0 product (L2-resident trunk kernel), 1 SIMT validation kernel, 3 one launch per layer pass, "
                "4 whole-frame persistent trunk kernel", cfg->conv_impl);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, NESR_E_CUDA, "no CUDA device (%s): libnesr_b200 has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, NESR_E_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
  DeviceGuard device_guard(cfg->device);
  if (!device_guard.ok) return fail(nullptr, NESR_E_CUDA, "cudaSetDevice(%d) failed", cfg->device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail(nullptr, NESR_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, NESR_E_CUDA, "device %d is sm_%d%d; libnesr_b200 is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
  nesr_b200_handle* h = new nesr_b200_handle();
  h->cfg = *cfg;
  h->num_sms = prop.multiProcessorCount;
  // evict_last cache hints only hold lines inside the persisting-L2 set-aside, which defaults to zero
  if (const char* ps = getenv("NESR_B200_L2_PERSIST_MB")) {
    size_t want = (size_t)atoi(ps) << 20;
    if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
    cudaError_t el = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    fprintf(stderr, "nesr_b200: persisting L2 set-aside %zu MB (max %d MB, L2 %d MB): %s\n", want >> 20,
            prop.persistingL2CacheMaxSize >> 20, prop.l2CacheSize >> 20, cudaGetErrorString(el));
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete h;
    return fail(nullptr, NESR_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  }
  h->encode = (EncodeTiledFn)fn;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess ||
      (e = cudaEventCreate(&h->evc0)) != cudaSuccess || (e = cudaEventCreate(&h->evc1)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_own, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming)) != cudaSuccess ||
      (e = conv3x3_fold_configure()) != cudaSuccess ||
      (e = conv3x3_body_configure()) != cudaSuccess || (e = conv3x3_trunk_configure()) != cudaSuccess || (e = cudaMalloc(&h->d_gbar, 2048 * 128)) != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    nesr_b200_destroy(h);
    return fail(nullptr, NESR_E_CUDA, "device setup failed: %s", msg.c_str());
  }
  build_layers(h);
#if NESR_PROF
  if (const char* dbg = getenv("NESR_B200_DEBUG_FLAGS")) h->debug_flags = atoi(dbg);   // timing experiments: -DNESR_PROF=1 builds only
#endif
  if (const char* pin = getenv("NESR_B200_L2_PIN")) h->l2_pin_chunks = atoi(pin);
  if (const char* sg = getenv("NESR_B200_SHARED_G")) h->shared_g = atoi(sg);
  if (const char* mp = getenv("NESR_B200_MAX_PIECES")) h->max_pieces = atoi(mp);
  if (const char* al = getenv("NESR_B200_ARENA_LIMIT_MB")) h->arena_limit = (size_t)atoll(al) << 20;
  *out = h;
  return NESR_OK;
}

int nesr_b200_destroy(nesr_b200_handle* h) {
  if (!h) return NESR_OK;
  DeviceGuard device_guard(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_batches(h);
  if (h->arena_base) cudaFree(h->arena_base);
  if (h->d_wpack) cudaFree(h->d_wpack);
  if (h->d_bias) cudaFree(h->d_bias);
  if (h->d_wfold) cudaFree(h->d_wfold);
  if (h->d_wmerge) cudaFree(h->d_wmerge);
  if (h->d_gbar) cudaFree(h->d_gbar);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->d_tmp) cudaFree(h->d_tmp);
  if (h->d_pre) cudaFree(h->d_pre);
  for (int i = 0; i < 2; ++i) if (h->d_nlm_w[i]) cudaFree(h->d_nlm_w[i]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (cudaEvent_t ev : h->ev_trunk) cudaEventDestroy(ev);
  if (h->evc0) cudaEventDestroy(h->evc0);
  if (h->evc1) cudaEventDestroy(h->evc1);
  if (h->ev_own) cudaEventDestroy(h->ev_own);
  if (h->ev_ext) cudaEventDestroy(h->ev_ext);
  for (cudaEvent_t ev : h->ev_group) cudaEventDestroy(ev);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return NESR_OK;
}

const char* nesr_b200_last_error(const nesr_b200_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int nesr_b200_load_weight(nesr_b200_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  if (!h || !name || !data || !shape) return fail(h, NESR_E_INVALID, "null argument");
  auto it = h->expect.find(name);
  if (it == h->expect.end()) return fail(h, NESR_E_WEIGHTS, "unexpected key in state_dict: %s", name);
  const std::vector<int64_t>& want = it->second;
  bool ok = (size_t)ndim == want.size();
  int64_t numel = 1;
  for (int i = 0; ok && i < ndim; ++i) { ok = shape[i] == want[i]; numel *= shape[i]; }
  if (!ok) return fail(h, NESR_E_WEIGHTS, "size mismatch for %s", name);
  h->staged[name].assign(data, data + numel);
  h->finalized = false;
  return NESR_OK;
}

int nesr_b200_finalize_weights(nesr_b200_handle* h) {
  if (!h) return NESR_E_INVALID;
  for (const auto& kv : h->expect)
    if (!h->staged.count(kv.first)) return fail(h, NESR_E_WEIGHTS, "missing key in state_dict: %s", kv.first.c_str());
  DEVICE_SCOPE(h);
  std::vector<uint16_t> arena((size_t)h->w_rows * 64, 0);
  std::vector<float> bias(h->layers.size() * 64, 0.f);
  for (const Layer& L : h->layers) {
    pack_layer(L, h->staged[L.name + ".weight"].data(), arena.data());
    const std::vector<float>& bv = h->staged[L.name + ".bias"];
    std::copy(bv.begin(), bv.end(), bias.begin() + L.bias_off);
  }
  std::vector<uint16_t> farena((size_t)h->wfold_rows * 64, 0);
  for (const Layer& L : h->layers)
    for (int ps = 0; ps < L.fold_passes; ++ps) pack_layer_fold(L, ps, h->staged[L.name + ".weight"].data(), farena.data());
  if (!h->d_wfold) CUDA_TRY(h, cudaMalloc(&h->d_wfold, farena.size() * 2));
  CUDA_TRY(h, cudaMemcpy(h->d_wfold, farena.data(), farena.size() * 2, cudaMemcpyHostToDevice));
  const int fbox[3] = {48, 96, 192};
  for (int i = 0; i < 3; ++i) {
    int rc = make_map(h, &h->m_wf[i], h->d_wfold, 64, h->wfold_rows, fbox[i]);
    if (rc) return rc;
  }
  {  // merged trunk weights: per dense block the eight sweep blocks of layout.h TrunkSweep
    const int nrdb = h->cfg.num_block * 3;
    h->wmerge_rows = (int64_t)nrdb * kMergeRowsPerBlock;
    std::vector<uint16_t> marena((size_t)h->wmerge_rows * 64, 0);
    for (int r = 0; r < nrdb; ++r) {
      const Layer* L[6];
      const float* W[6];
      for (int k = 1; k <= 5; ++k) { L[k] = &h->layers[1 + 5 * r + (k - 1)]; W[k] = h->staged[L[k]->name + ".weight"].data(); }
      uint16_t* base = marena.data() + (int64_t)r * kMergeRowsPerBlock * 64;
      const int fmt = h->cfg.body_format;
      pack_sweep(L[1], L[2], W[1], W[2], 0, fmt, base + (int64_t)kSweepRowOff[0] * 64);       // S1 x      -> conv1 | conv2
      pack_sweep(nullptr, L[2], nullptr, W[2], 1, fmt, base + (int64_t)kSweepRowOff[1] * 64); // S2 x1     -> conv2
      pack_sweep(L[3], L[4], W[3], W[4], 0, fmt, base + (int64_t)kSweepRowOff[2] * 64);       // S3 x      -> conv3 | conv4
      pack_sweep(L[3], L[4], W[3], W[4], 1, fmt, base + (int64_t)kSweepRowOff[3] * 64);       // S4 x1, x2 -> conv3 | conv4
      pack_sweep(nullptr, L[4], nullptr, W[4], 2, fmt, base + (int64_t)kSweepRowOff[4] * 64); // S5 x3     -> conv4
      for (int c = 0; c < 3; ++c)                                                               // S6-S8 x | x1,x2 | x3,x4 -> conv5
        pack_sweep(L[5], L[5], W[5], W[5], c, fmt, base + (int64_t)kSweepRowOff[5 + c] * 64);
    }
    if (h->d_wmerge) { cudaFree(h->d_wmerge); h->d_wmerge = nullptr; }
    CUDA_TRY(h, cudaMalloc(&h->d_wmerge, marena.size() * 2));
    CUDA_TRY(h, cudaMemcpy(h->d_wmerge, marena.data(), marena.size() * 2, cudaMemcpyHostToDevice));
    const int mbox[2] = {192, 160};
    for (int i = 0; i < 2; ++i) {
      int rc = make_map(h, &h->m_wm[i], h->d_wmerge, 64, h->wmerge_rows, mbox[i]);
      if (rc) return rc;
    }
  }
  if (!h->d_wpack) CUDA_TRY(h, cudaMalloc(&h->d_wpack, arena.size() * 2));
  if (!h->d_bias) CUDA_TRY(h, cudaMalloc(&h->d_bias, bias.size() * 4));
  CUDA_TRY(h, cudaMemcpy(h->d_wpack, arena.data(), arena.size() * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  const int box[3] = {16, 32, 64};
  for (int i = 0; i < 3; ++i) {
    int rc = make_map(h, &h->m_w[i], h->d_wpack, 64, h->w_rows, box[i]);
    if (rc) return rc;
  }
  // pageable synchronous copies may return before their DMA has landed, and the kernels run on non-blocking streams that
  // have no implicit ordering with the legacy stream: make the uploads globally complete here
  CUDA_TRY(h, cudaDeviceSynchronize());
  h->finalized = true;
  return NESR_OK;
}

int nesr_b200_enhance_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W, int64_t in_stride, int32_t tile,
                         int32_t tile_pad, int32_t pre_pad, uint8_t* out_bgr, int64_t out_stride, int32_t flags) {
  return enhance_impl(h, in_bgr, 1, H, W, in_stride, 0, tile, tile_pad, pre_pad, 0, 0, 1, out_bgr, out_stride, 0, flags);
}

int nesr_b200_enhance_batch_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t n_frames, int32_t H, int32_t W,
                               int64_t in_stride, int64_t in_frame_stride, int32_t tile, int32_t tile_pad, int32_t pre_pad,
                               uint8_t* out_bgr, int64_t out_stride, int64_t out_frame_stride, int32_t flags) {
  return enhance_impl(h, in_bgr, n_frames, H, W, in_stride, in_frame_stride, tile, tile_pad, pre_pad, 0, 0, 1, out_bgr,
                      out_stride, out_frame_stride, flags);
}

int nesr_b200_tile_count(int32_t H, int32_t W, int32_t tile, int32_t pre_pad, int32_t scale) {
  if (H < 1 || W < 1 || tile < 0 || pre_pad < 0) return NESR_E_INVALID;
  const Grid g = tile_grid_dims(H, W, tile, pre_pad, scale);
  return g.tiles_x * g.tiles_y;
}

// Host-only: builds the tile-group plan a call with these arguments would use (no CUDA calls, works without a GPU) and
// verifies its invariants.  out[]: 0 groups, 1 tiles, 2 feature pixels, 3 strip rows of level 0 (all groups), 4 largest
// number of output rows owned by one CTA, 5 groups the TMEM-resident trunk kernel can run, 6 reserved (0),
// 7 halo rows (2 per band, level 0).  Returns NESR_OK, or NESR_E_STATE with the violated invariant in last_error(NULL).
int nesr_b200_debug_plan(int32_t n_frames, int32_t H, int32_t W, int32_t tile, int32_t tile_pad, int32_t pre_pad, int32_t num_sms,
                         int32_t conv_impl, int64_t max_batch_pixels, int32_t pairs, int32_t sets, int64_t* out) {
  if (!out || n_frames < 1 || H < 2 || W < 2 || num_sms < 1) return fail(nullptr, NESR_E_INVALID, "bad arguments");
  nesr_b200_handle hd;
  nesr_b200_default_config(&hd.cfg, 0);
  hd.cfg.conv_impl = conv_impl; hd.cfg.max_batch_pixels = max_batch_pixels;
  hd.num_sms = num_sms;
  if (const char* mp = getenv("NESR_B200_MAX_PIECES")) hd.max_pieces = atoi(mp);
  (void)pairs; (void)sets;                                     // reserved (round-1 planner variants, removed): ignored
  PlanKey key; key.n_frames = n_frames; key.H = H; key.W = W; key.tile = tile; key.tile_pad = tile_pad; key.pre_pad = pre_pad;
  key.first = 0; key.count = 0; key.whole = 1;
  const int rc = plan_groups(&hd, key, H * hd.cfg.scale, W * hd.cfg.scale);
  if (rc) return fail(nullptr, rc, "%s", hd.error.c_str());
  for (int k = 0; k < 8; ++k) out[k] = 0;
  out[0] = (int64_t)hd.batches.size();
  for (size_t bi = 0; bi < hd.batches.size(); ++bi) {
    const Batch& b = hd.batches[bi];
    out[1] += (int64_t)b.tiles.size();
    out[5] += b.trunk_fits;
    for (int level = 0; level < 3; ++level) {
      const LevelPlan& lp = b.lv[level];
      // every pixel of every tile is owned by exactly one (band, segment, lane); nothing outside a tile is owned
      std::vector<std::vector<uint8_t>> cover(b.tiles.size());
      for (size_t t = 0; t < b.tiles.size(); ++t) cover[t].assign((size_t)b.tiles[t].lv[level].h * b.tiles[t].lv[level].w, 0);
      if ((int)lp.cta_off.size() != lp.fold_grid + 1 || lp.fold_grid > num_sms)
        return fail(nullptr, NESR_E_STATE, "group %zu level %d: grid %d, %zu offsets", bi, level, lp.fold_grid, lp.cta_off.size());
      for (int c = 0; c < lp.fold_grid; ++c) {
        int rows = 0;
        for (int q = lp.cta_off[c]; q < lp.cta_off[c + 1]; ++q) {
          const FoldBand& band = lp.bands[q];
          if (band.rows < 1 || band.nseg < 1 || band.nseg > kMaxFoldSegs)
            return fail(nullptr, NESR_E_STATE, "group %zu level %d: band %d has %d rows, %d segments", bi, level, q, band.rows, band.nseg);
          rows += band.rows;
          if (level == 0) { out[3] += band.rows; out[7] += 2; }
          int lanes = 0;
          for (int sgi = 0; sgi < band.nseg; ++sgi) {
            const FoldSeg& sg = lp.segs[band.seg0 + sgi];
            const LevelGeom& g = b.tiles[sg.tile].lv[level];
            const int halo = 2 + (level > 0 ? 1 : 0);      // levels above 0 may be read up-sampled: slab rows start at the even pixel x0 - 2
            if (sg.lane0 < lanes || (sg.lane0 & 7) || sg.lane0 + sg.width > kBlockPixels || sg.lane0 + sg.width + halo > 136 ||
                sg.x0 < 0 || sg.x0 + sg.width > g.w || sg.y0 < 0 || sg.y0 + sg.h > g.h || (level > 0 && (sg.x0 & 1)))
              return fail(nullptr, NESR_E_STATE, "group %zu level %d: segment %d out of range", bi, level, band.seg0 + sgi);
            lanes = sg.lane0 + (sg.width + halo + 7) / 8 * 8;
            for (int r = band.r0; r < band.r0 + band.rows && r < sg.h; ++r)
              for (int x = sg.x0; x < sg.x0 + sg.width; ++x)
                if (++cover[sg.tile][(size_t)(sg.y0 + r) * g.w + x] != 1)
                  return fail(nullptr, NESR_E_STATE, "group %zu level %d: tile %d pixel (%d,%d) owned twice", bi, level, sg.tile, sg.y0 + r, x);
          }
        }
        if (level == 0) out[4] = std::max<int64_t>(out[4], rows);
      }
      for (size_t t = 0; t < b.tiles.size(); ++t)
        for (size_t i2 = 0; i2 < cover[t].size(); ++i2)
          if (cover[t][i2] != 1) return fail(nullptr, NESR_E_STATE, "group %zu level %d: tile %zu pixel %zu not owned", bi, level, t, i2);
      if (level == 0) {
        for (const TileGeom& t : b.tiles) out[2] += (int64_t)t.lv[0].h * t.lv[0].w;
        if (b.trunk_fits) {
          if (!trunk_schedule_fits(lp)) return fail(nullptr, NESR_E_STATE, "group %zu: trunk fit flag wrong", bi);
          // Row-dependency table, checked pixel by pixel against an owner map built independently of build_trunk_deps: every
          // pixel of every slab row that feeds an output row must be covered by need[c][t][lane of its owner] >= the number of
          // rows its owner has completed when it completes that pixel's row.
          if ((int)lp.deps.size() != lp.fold_grid * kTrunkMaxDeps || (int)lp.need.size() != lp.fold_grid * kTrunkMaxSlabRows * kTrunkMaxDeps)
            return fail(nullptr, NESR_E_STATE, "group %zu: dependency table size", bi);
          std::vector<std::vector<int32_t>> owner(b.tiles.size());                  // cta * 32 + slot per pixel
          for (size_t t = 0; t < b.tiles.size(); ++t) owner[t].assign((size_t)b.tiles[t].lv[0].h * b.tiles[t].lv[0].w, -1);
          for (int c = 0; c < lp.fold_grid; ++c) {
            int slot0 = 0;
            for (int q = lp.cta_off[c]; q < lp.cta_off[c + 1]; ++q) {
              const FoldBand& band = lp.bands[q];
              for (int sgi = 0; sgi < band.nseg; ++sgi) {
                const FoldSeg& sg = lp.segs[band.seg0 + sgi];
                const LevelGeom& g = b.tiles[sg.tile].lv[0];
                for (int r = band.r0; r < band.r0 + band.rows && r < sg.h; ++r)
                  for (int x = sg.x0; x < sg.x0 + sg.width; ++x) owner[sg.tile][(size_t)(sg.y0 + r) * g.w + x] = c * 32 + slot0 + (r - band.r0);
              }
              slot0 += band.rows;
            }
          }
          // the order in which every CTA sweeps its slab rows and completes its output rows: every slab row once, every output
          // row once and only after its three slab rows; the pieces of a packed band agree on the tile row modulo 8
          std::vector<TrunkOrder> order(lp.fold_grid);
          std::vector<std::array<uint8_t, kTrunkMaxRows>> done_at(lp.fold_grid);
          for (int c = 0; c < lp.fold_grid; ++c) {
            TrunkOrder& o = order[c];
            if (!cta_order(lp, c, o)) return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d has no phase order", bi, c);
            int nin = 0, nout = 0;
            uint32_t swept[kTrunkMaxBands] = {0, 0, 0, 0}, outs = 0;
            for (int q = lp.cta_off[c]; q < lp.cta_off[c + 1]; ++q) {
              const FoldBand& band = lp.bands[q];
              nin += band.rows + 2; nout += band.rows;
              int res = -1;
              for (int sgi = 0; sgi < band.nseg; ++sgi) {
                const FoldSeg& sg = lp.segs[band.seg0 + sgi];
                if (band.r0 >= sg.h) continue;
                if (res >= 0 && res != ((sg.y0 + band.r0) & 7)) {
                  if (getenv("NESR_B200_PLAN_DEBUG"))
                    for (int k2 = 0; k2 < band.nseg; ++k2) {
                      const FoldSeg& s2 = lp.segs[band.seg0 + k2];
                      fprintf(stderr, "  band r0 %d rows %d: seg %d tile %d x0 %d width %d lane0 %d y0 %d h %d\n", band.r0, band.rows, k2, s2.tile, s2.x0, s2.width, s2.lane0, s2.y0, s2.h);
                    }
                  return fail(nullptr, NESR_E_STATE, "group %zu: packed pieces out of phase", bi);
                }
                res = (sg.y0 + band.r0) & 7;
              }
            }
            if (o.n_in != nin || o.n_out != nout) return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d phase order size", bi, c);
            done_at[c].fill(0);
            int t_out = 0;
            for (int t = 0; t < o.n_in; ++t) {
              const int bb = o.in_band[t], i = o.in_row[t];
              if (bb >= lp.cta_off[c + 1] - lp.cta_off[c] || i > lp.bands[lp.cta_off[c] + bb].rows + 1 || ((swept[bb] >> i) & 1))
                return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d slab row swept twice", bi, c);
              swept[bb] |= 1u << i;
              int slot0 = 0;
              for (int q = 0; q < bb; ++q) slot0 += lp.bands[lp.cta_off[c] + q].rows;
              // rows listed as complete after this slab row must have all three of theirs
              while (t_out < o.n_out) {
                const int sl = o.out_slot[t_out], j = sl - slot0;
                if (j < 0 || j >= lp.bands[lp.cta_off[c] + bb].rows || ((swept[bb] >> j) & 7u) != 7u || j < i - 2 || j > i) break;
                if ((outs >> sl) & 1) return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d output row completed twice", bi, c);
                outs |= 1u << sl;
                done_at[c][sl] = (uint8_t)++t_out;
              }
            }
            if (t_out != nout) return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d completes %d of %d rows", bi, c, t_out, nout);
          }
          for (int c = 0; c < lp.fold_grid; ++c) {
            const int32_t* dep = &lp.deps[(size_t)c * kTrunkMaxDeps];
            if (dep[0] != c) return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d is not its own first dependency", bi, c);
            for (int t = 0; t < order[c].n_in; ++t) {
              {
                const FoldBand& band = lp.bands[lp.cta_off[c] + order[c].in_band[t]];
                const int i = order[c].in_row[t] - 1;
                for (int sgi = 0; sgi < band.nseg; ++sgi) {
                  const FoldSeg& sg = lp.segs[band.seg0 + sgi];
                  const LevelGeom& g = b.tiles[sg.tile].lv[0];
                  const int last_out = std::min(band.r0 + band.rows, sg.h) - 1;        // last output strip row of this piece in the band
                  if (last_out < band.r0) continue;
                  const int srow = band.r0 + i, y = sg.y0 + srow;
                  if (srow > last_out + 1 || y < 0 || y >= g.h) continue;
                  for (int x = std::max(sg.x0 - 1, 0); x <= std::min(sg.x0 + sg.width, g.w - 1); ++x) {
                    const int32_t o = owner[sg.tile][(size_t)y * g.w + x];
                    if (o < 0) return fail(nullptr, NESR_E_STATE, "group %zu: unowned pixel in a slab row", bi);
                    int lane = -1;
                    for (int k = 0; k < kTrunkMaxDeps; ++k) if (dep[k] == o / 32) { lane = k; break; }
                    if (lane < 0 || lp.need[((size_t)c * kTrunkMaxSlabRows + t) * kTrunkMaxDeps + lane] < done_at[o / 32][o % 32])
                      return fail(nullptr, NESR_E_STATE, "group %zu: CTA %d slab row %d does not wait for CTA %d row %d", bi, c, t, o / 32, o % 32);
                  }
                }
              }
            }
            for (int tt = 0; tt < kTrunkMaxSlabRows; ++tt)
              for (int k = 0; k < kTrunkMaxDeps; ++k)
                if (lp.need[((size_t)c * kTrunkMaxSlabRows + tt) * kTrunkMaxDeps + k] > kTrunkMaxRows)
                  return fail(nullptr, NESR_E_STATE, "group %zu: row requirement out of range", bi);
          }
        }
      }
    }
  }
  free_batches(&hd);
  return NESR_OK;
}

int nesr_b200_enhance_tiles_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W, int64_t in_stride, int32_t tile,
                               int32_t tile_pad, int32_t pre_pad, int32_t tile_first, int32_t tile_count, uint8_t* out_bgr,
                               int64_t out_stride, int32_t flags) {
  if (tile_count == 0) return NESR_OK;
  return enhance_impl(h, in_bgr, 1, H, W, in_stride, 0, tile, tile_pad, pre_pad, tile_first, tile_count, 0, out_bgr, out_stride, 0, flags);
}

int nesr_b200_enhance_tiles_packed_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W, int64_t in_stride, int32_t tile,
                                      int32_t tile_pad, int32_t pre_pad, int32_t tile_first, int32_t tile_count, uint8_t* slots,
                                      int32_t slot_w, int32_t slot_h, int32_t flags) {
  if (tile_count == 0) return NESR_OK;
  if (!h) return NESR_E_INVALID;
  const int s = h->cfg.scale;
  const int need_w = (tile > 0 ? std::min(tile, W) : W) * s, need_h = (tile > 0 ? std::min(tile, H) : H) * s;
  if (slot_w < need_w || slot_h < need_h) return fail(h, NESR_E_INVALID, "tile slot %dx%d smaller than a tile's output %dx%d", slot_w, slot_h, need_w, need_h);
  return enhance_impl(h, in_bgr, 1, H, W, in_stride, 0, tile, tile_pad, pre_pad, tile_first, tile_count, 0, slots, (int64_t)slot_w * 3,
                      (int64_t)slot_w * 3 * slot_h, flags | NESR_PTR_OUT_DEVICE, 1);
}

int nesr_b200_enhance_tile_list_packed_u8(nesr_b200_handle* h, const uint8_t* in_bgr, int32_t H, int32_t W, int64_t in_stride, int32_t tile,
                                          int32_t tile_pad, int32_t pre_pad, const int32_t* tile_ids, int32_t tile_count, uint8_t* slots,
                                          int32_t slot_w, int32_t slot_h, int32_t flags) {
  if (tile_count == 0) return NESR_OK;
  if (!h) return NESR_E_INVALID;
  if (!tile_ids || tile_count < 0) return fail(h, NESR_E_INVALID, "tile list: bad arguments");
  const int s = h->cfg.scale;
  const int need_w = (tile > 0 ? std::min(tile, W) : W) * s, need_h = (tile > 0 ? std::min(tile, H) : H) * s;
  if (slot_w < need_w || slot_h < need_h) return fail(h, NESR_E_INVALID, "tile slot %dx%d smaller than a tile's output %dx%d", slot_w, slot_h, need_w, need_h);
  return enhance_impl(h, in_bgr, 1, H, W, in_stride, 0, tile, tile_pad, pre_pad, 0, tile_count, 0, slots, (int64_t)slot_w * 3,
                      (int64_t)slot_w * 3 * slot_h, flags | NESR_PTR_OUT_DEVICE, 1, tile_ids);
}

int nesr_b200_enhance_head_u8(nesr_b200_handle* h, const uint8_t* in_rgb, int32_t H, int32_t W, int64_t in_stride, int32_t force_3channel,
                              uint8_t* out_rgb, int64_t out_stride, int32_t flags) {
  if (!h) return NESR_E_INVALID;
  if (H < 1 || W < 1) return fail(h, NESR_E_INVALID, "bad image arguments (H=%d W=%d)", H, W);
  // the 12-channel scale-4 architecture is the x2plus network behind its un-shuffle: the H x W image's 12 channels are the un-shuffle of a
  // virtual 2H x 2W frame, the output is 4H x 4W; one untiled forward (the reference tiles with its own host loop above this call)
  return enhance_impl(h, in_rgb, 1, 2 * H, 2 * W, in_stride, 0, 0, 0, 0, 0, 0, 1, out_rgb, out_stride, 0, flags, 0, nullptr, force_3channel ? 2 : 1);
}

namespace {
int unpack_impl(nesr_b200_handle* h, const uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t H, int32_t W, int32_t tile, int32_t pre_pad,
                int32_t tile_first, int32_t tile_count, const int32_t* tile_ids, uint8_t* out_bgr, int64_t out_stride);
}

int nesr_b200_unpack_tiles_u8(nesr_b200_handle* h, const uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t H, int32_t W, int32_t tile,
                              int32_t pre_pad, int32_t tile_first, int32_t tile_count, uint8_t* out_bgr, int64_t out_stride) {
  return unpack_impl(h, slots, slot_w, slot_h, H, W, tile, pre_pad, tile_first, tile_count, nullptr, out_bgr, out_stride);
}

int nesr_b200_unpack_tile_list_u8(nesr_b200_handle* h, const uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t H, int32_t W, int32_t tile,
                                  int32_t pre_pad, const int32_t* tile_ids, int32_t slot_count, uint8_t* out_bgr, int64_t out_stride) {
  if (!tile_ids && slot_count > 0) return fail(h, NESR_E_INVALID, "unpack_tile_list: null list");
  return unpack_impl(h, slots, slot_w, slot_h, H, W, tile, pre_pad, 0, slot_count, tile_ids, out_bgr, out_stride);
}

namespace {
int unpack_impl(nesr_b200_handle* h, const uint8_t* slots, int32_t slot_w, int32_t slot_h, int32_t H, int32_t W, int32_t tile, int32_t pre_pad,
                int32_t tile_first, int32_t tile_count, const int32_t* tile_ids, uint8_t* out_bgr, int64_t out_stride) {
  if (tile_count == 0) return NESR_OK;
  if (!h) return NESR_E_INVALID;
  if (!slots || !out_bgr || H < 1 || W < 1 || tile < 0 || tile_first < 0 || tile_count < 0) return fail(h, NESR_E_INVALID, "unpack_tiles: bad arguments");
  DEVICE_SCOPE(h);
  if (int wrc = wait_external(h)) return wrc;
  const int s = h->cfg.scale;
  const Grid g = tile_grid_dims(H, W, tile, pre_pad, s);
  if (!tile_ids && tile_first + tile_count > g.tiles_x * g.tiles_y) return fail(h, NESR_E_INVALID, "unpack_tiles: tile range outside the grid");
  if (tile_ids)
    for (int k = 0; k < tile_count; ++k)
      if (tile_ids[k] >= g.tiles_x * g.tiles_y) return fail(h, NESR_E_INVALID, "unpack_tile_list: tile %d outside the grid", tile_ids[k]);
  cudaError_t e = launch_unpack_tiles(slots, slot_w, slot_h, g.tiles_x, (tile > 0 ? tile : std::max(H, W)) * s, H * s, W * s, tile_first, tile_count, tile_ids,
                                      out_bgr, out_stride, h->stream);
  h->stats.kernel_launches++;
  if (e != cudaSuccess) return fail(h, NESR_E_CUDA, "unpack_tiles launch failed: %s", cudaGetErrorString(e));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return NESR_OK;
}
}  // namespace

namespace {
int forward_nchw(nesr_b200_handle* h, const float* x, bool unshuffled, int32_t n, int32_t H, int32_t W, float* y, void* stream) {
  if (!h) return NESR_E_INVALID;
  if (!unshuffled && h->cfg.feat_in_ch > 0 && h->cfg.feat_in_ch != 12)
    return fail(h, NESR_E_STATE, "this handle's conv_first takes %d feature channels: use nesr_b200_forward_feat_f32", h->cfg.feat_in_ch);
  if (!h->finalized) return fail(h, NESR_E_STATE, "weights not finalized");
  if (!x || !y || n < 1 || H < 2 || W < 2) return fail(h, NESR_E_INVALID, "bad tensor arguments");
  if ((H | W) & 1) return fail(h, NESR_E_INVALID, "pixel_unshuffle(2): H and W must be even (got %dx%d)", H, W);
  DEVICE_SCOPE(h);
  cudaStream_t s = (cudaStream_t)stream;      // literally the caller's stream; NULL is the legacy default stream
  const int sc = h->cfg.scale;
  PlanKey key; key.n_frames = n; key.H = H; key.W = W; key.tile = 0; key.whole = 1;
  int rc = build_plan(h, key, H * sc, W * sc);
  if (rc) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev_own, h->stream));          // plan uploads / memset (and any earlier u8 call) ran on our own stream
  CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_own, 0));
  PackParams pk{};
  if (unshuffled) { pk.in_f32_12 = x; pk.feat_ch = h->cfg.feat_in_ch > 0 ? h->cfg.feat_in_ch : 12; } else pk.in_f32 = x;
  pk.H = H; pk.W = W; pk.pre_pad = 0;
  Sink sink; sink.out_f32 = y; sink.out_h = H * sc; sink.out_w = W * sc;
  for (const Batch& b : h->batches)
    if ((rc = forward_batch(h, b, pk, sink, s, false, false))) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev_ext, s));
  h->ext_pending = true;
  return NESR_OK;
}
}  // namespace

int nesr_b200_forward_nchw_f32(nesr_b200_handle* h, const float* x, int32_t n, int32_t H, int32_t W, float* y, void* stream) {
  return forward_nchw(h, x, false, n, H, W, y, stream);
}

int nesr_b200_forward_nchw12_f32(nesr_b200_handle* h, const float* x12, int32_t n, int32_t H, int32_t W, float* y, void* stream) {
  if (!h) return NESR_E_INVALID;
  if (H < 1 || W < 1) return fail(h, NESR_E_INVALID, "bad tensor arguments");
  if (h->cfg.feat_in_ch > 0 && h->cfg.feat_in_ch != 12) return fail(h, NESR_E_STATE, "this handle's conv_first takes %d feature channels, not 12", h->cfg.feat_in_ch);
  return forward_nchw(h, x12, true, n, 2 * H, 2 * W, y, stream);        // the 12 channels are the un-shuffle of a 2H x 2W image
}

int nesr_b200_forward_feat_f32(nesr_b200_handle* h, const float* feat, int32_t n, int32_t C, int32_t H, int32_t W, float* y, void* stream) {
  if (!h) return NESR_E_INVALID;
  if (H < 1 || W < 1) return fail(h, NESR_E_INVALID, "bad tensor arguments");
  const int want = h->cfg.feat_in_ch > 0 ? h->cfg.feat_in_ch : 12;
  if (C != want) return fail(h, NESR_E_INVALID, "feature tensor has %d channels, this handle's conv_first takes %d", C, want);
  return forward_nchw(h, feat, true, n, 2 * H, 2 * W, y, stream);       // the engine's geometry is that of a 2H x 2W x2plus input
}

int nesr_b200_blend_u8(nesr_b200_handle* h, const uint8_t* const* members, int32_t K, int32_t H, int32_t W, const double* weights,
                       uint8_t* out, int32_t flags) {
  if (!h) return NESR_E_INVALID;
  if (!members || !out || K < 2 || K > kMaxBlendMembers || H < 1 || W < 1) return fail(h, NESR_E_INVALID, "blend: need 2..%d members", kMaxBlendMembers);
  DEVICE_SCOPE(h);
  if (int wrc = wait_external(h)) return wrc;
  const int64_t n = (int64_t)H * W * 3;
  BlendParams p{};
  p.k = K; p.nbytes = n;
  for (int i = 0; i < K; ++i) p.weights[i] = weights ? weights[i] : 1.0 / K;
  int rc;
  if (!(flags & NESR_PTR_IN_DEVICE)) {
    const size_t slot = (size_t)round_up(n, 256);
    if ((rc = ensure(h, &h->d_in, &h->d_in_bytes, slot * K))) return rc;
    for (int i = 0; i < K; ++i) {
      CUDA_TRY(h, cudaMemcpyAsync(h->d_in + slot * i, members[i], n, cudaMemcpyHostToDevice, h->stream));
      p.members[i] = h->d_in + slot * i;
    }
  } else {
    for (int i = 0; i < K; ++i) p.members[i] = members[i];
  }
  if (!(flags & NESR_PTR_OUT_DEVICE)) {
    if ((rc = ensure(h, &h->d_out, &h->d_out_bytes, (size_t)n))) return rc;
    p.out = h->d_out;
  } else {
    p.out = out;
  }
  cudaEventRecord(h->ev0, h->stream);
  cudaError_t e = launch_blend(p, h->stream);
  cudaEventRecord(h->ev1, h->stream);
  h->stats.kernel_launches++;
  if (e != cudaSuccess) return fail(h, NESR_E_CUDA, "blend launch failed: %s", cudaGetErrorString(e));
  if (!(flags & NESR_PTR_OUT_DEVICE)) CUDA_TRY(h, cudaMemcpyAsync(out, h->d_out, n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.last_device_ms = ms;
  return NESR_OK;
}

int nesr_b200_preprocess_u8(nesr_b200_handle* h, const uint8_t* rgb, int32_t H, int32_t W, float denoise_h, float denoise_h_color,
                            float clahe_clip, int32_t tiles_x, int32_t tiles_y, uint8_t* out, int32_t flags) {
  if (!h) return NESR_E_INVALID;
  if (!rgb || !out || H < 1 || W < 1 || tiles_x < 1 || tiles_y < 1 || tiles_x * tiles_y > 4096)
    return fail(h, NESR_E_INVALID, "preprocess: bad arguments");
  if ((int64_t)H * W > (int64_t)1 << 30) return fail(h, NESR_E_INVALID, "preprocess: image too large");
  DEVICE_SCOPE(h);
  if (int wrc = wait_external(h)) return wrc;
  const int64_t n = (int64_t)H * W * 3;
  const uint8_t* d_in = rgb;
  uint8_t* d_out = out;
  int rc;
  if (!(flags & NESR_PTR_IN_DEVICE)) {
    if ((rc = ensure(h, &h->d_in, &h->d_in_bytes, (size_t)n))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_in, rgb, n, cudaMemcpyHostToDevice, h->stream));
    d_in = h->d_in;
  }
  if (!(flags & NESR_PTR_OUT_DEVICE)) {
    if ((rc = ensure(h, &h->d_out, &h->d_out_bytes, (size_t)n))) return rc;
    d_out = h->d_out;
  }
  if ((rc = ensure(h, &h->d_pre, &h->d_pre_bytes, preprocess_workspace_bytes(H, W, tiles_x, tiles_y)))) return rc;
  const bool denoise = denoise_h > 0.f;                  // the reference skips the denoiser at level 0 (nesr/nesr.py:671)
  if (denoise) {
    const float hs[2] = {denoise_h, denoise_h_color > 0.f ? denoise_h_color : denoise_h};
    for (int i = 0; i < 2; ++i) {
      if (h->nlm_h[i] == hs[i] && h->d_nlm_w[i]) continue;
      const std::vector<int32_t> tab = nlm_weight_table(hs[i], i + 1);
      if (h->d_nlm_w[i]) { cudaFree(h->d_nlm_w[i]); h->d_nlm_w[i] = nullptr; }
      CUDA_TRY(h, cudaMalloc(&h->d_nlm_w[i], tab.size() * sizeof(int32_t)));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_nlm_w[i], tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));     // tab is a local
      h->n_nlm_w[i] = (int)tab.size();
      h->nlm_h[i] = hs[i];
    }
  }
  int launches = 0;
  cudaEventRecord(h->ev0, h->stream);
  cudaError_t e = launch_preprocess(d_in, d_out, H, W, denoise ? h->d_nlm_w[0] : nullptr, h->n_nlm_w[0], denoise ? h->d_nlm_w[1] : nullptr,
                                    h->n_nlm_w[1], clahe_clip, tiles_x, tiles_y, h->d_pre, &launches, h->stream);
  cudaEventRecord(h->ev1, h->stream);
  h->stats.kernel_launches += launches;
  if (e != cudaSuccess) return fail(h, NESR_E_CUDA, "preprocess launch failed: %s", cudaGetErrorString(e));
  if (!(flags & NESR_PTR_OUT_DEVICE)) CUDA_TRY(h, cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.last_device_ms = ms;
  return NESR_OK;
}

int nesr_b200_debug_lab_table(int32_t which, void* out, int32_t capacity_bytes) {
  int count = 0, elem = 0;
  const void* src = lab_table_host(which, &count, &elem);
  if (!src) return -1;
  if (out) {
    if (capacity_bytes < count * elem) return -1;
    memcpy(out, src, (size_t)count * elem);
  }
  return count * elem;
}

int nesr_b200_debug_nlm_weights(float h, int32_t channels, int32_t* out, int32_t capacity) {
  if (channels < 1 || channels > 2) return -1;
  const std::vector<int32_t> tab = nlm_weight_table(h, channels);
  if (out) {
    if (capacity < (int32_t)tab.size()) return -1;
    memcpy(out, tab.data(), tab.size() * sizeof(int32_t));
  }
  return (int32_t)tab.size();
}

}  // extern "C"
namespace {
// _postprocess_image (mask == null) and the segmentation-masked unsharp of _segment_and_enhance (mask: H x W u8 on the side `in` is on)
int sharpen_impl(nesr_b200_handle* h, const uint8_t* in, const uint8_t* mask, int32_t H, int32_t W, int32_t bgr, uint8_t* out, int32_t flags) {
  DEVICE_SCOPE(h);
  if (int wrc = wait_external(h)) return wrc;
  const int64_t n = (int64_t)H * W * 3;
  const uint8_t* d_in = in;
  uint8_t* d_out = out;
  int rc;
  if (!(flags & NESR_PTR_IN_DEVICE)) {
    if ((rc = ensure(h, &h->d_in, &h->d_in_bytes, (size_t)n))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_in, in, n, cudaMemcpyHostToDevice, h->stream));
    d_in = h->d_in;
  }
  if (!(flags & NESR_PTR_OUT_DEVICE)) {
    if ((rc = ensure(h, &h->d_out, &h->d_out_bytes, (size_t)n))) return rc;
    d_out = h->d_out;
  }
  if (d_in == d_out) return fail(h, NESR_E_INVALID, "sharpen: in-place operation is not supported");
  const uint8_t* d_mask = mask;
  if (mask && !(flags & NESR_PTR_IN_DEVICE)) {
    if ((rc = ensure(h, &h->d_tmp, &h->d_tmp_bytes, (size_t)H * W))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_tmp, mask, (size_t)H * W, cudaMemcpyHostToDevice, h->stream));
    d_mask = h->d_tmp;
  }
  cudaEventRecord(h->ev0, h->stream);
  cudaError_t e = launch_sharpen(d_in, d_out, H, W, bgr, d_mask, h->stream);
  cudaEventRecord(h->ev1, h->stream);
  h->stats.kernel_launches++;
  if (e != cudaSuccess) return fail(h, NESR_E_CUDA, "sharpen launch failed: %s", cudaGetErrorString(e));
  if (!(flags & NESR_PTR_OUT_DEVICE)) CUDA_TRY(h, cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.last_device_ms = ms;
  return NESR_OK;
}
}  // namespace
extern "C" {

int nesr_b200_sharpen_u8(nesr_b200_handle* h, const uint8_t* in, int32_t H, int32_t W, int32_t bgr, uint8_t* out, int32_t flags) {
  if (!h) return NESR_E_INVALID;
  if (!in || !out || H < 1 || W < 1) return fail(h, NESR_E_INVALID, "sharpen: bad arguments");
  return sharpen_impl(h, in, nullptr, H, W, bgr, out, flags);
}

int nesr_b200_masked_unsharp_u8(nesr_b200_handle* h, const uint8_t* in, const uint8_t* object_mask, int32_t H, int32_t W, int32_t bgr,
                                uint8_t* out, int32_t flags) {
  if (!h) return NESR_E_INVALID;
  if (!in || !object_mask || !out || H < 1 || W < 1) return fail(h, NESR_E_INVALID, "masked_unsharp: bad arguments");
  return sharpen_impl(h, in, object_mask, H, W, bgr, out, flags);
}

int nesr_b200_get_stats(const nesr_b200_handle* h, nesr_b200_stats* out) {
  if (!h || !out) return NESR_E_INVALID;
  *out = h->stats;
  return NESR_OK;
}

int nesr_b200_synchronize(nesr_b200_handle* h) {
  if (!h) return NESR_E_INVALID;
  DEVICE_SCOPE(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return NESR_OK;
}

// One conv layer on a single H x W "tile" (level-0 geometry), fp32 NCHW in/out on the host; operands
// are rounded to `fmt`, accumulation is fp32, the result is returned un-rounded (dst32a path).
int nesr_b200_debug_conv(nesr_b200_handle* h, int32_t impl, int32_t fmt, int32_t H, int32_t W, int32_t cin, int32_t cout,
                         const float* weight_oihw, const float* bias, const float* x_nchw, int32_t lrelu, float* y_nchw) {
  if (!h) return NESR_E_INVALID;
  if (!weight_oihw || !bias || !x_nchw || !y_nchw || H < 1 || W < 1 || cin < 1 || cin > kDense || cout < 1 || cout > 64 || (fmt & ~1) || (impl != 0 && impl != 1))
    return fail(h, NESR_E_INVALID, "debug_conv: bad arguments");
  DEVICE_SCOPE(h);
  Layer L;
  L.name = "debug"; L.cin = cin; L.cout = cout; L.fmt = fmt;
  L.cin16 = (int)round_up(cin, 16); L.nchunk = (L.cin16 + 63) / 64; L.npad = cout <= 16 ? 16 : (cout <= 32 ? 32 : 64);
  L.w_row0 = 0; L.bias_off = 0;
  const bool fits = conv3x3_fold_fits(L.cin16, L.npad);
  L.fold_passes = fits ? 1 : 2;
  L.fold_npad = fits ? L.npad : L.npad / 2;
  const int64_t pass_rows = (int64_t)3 * L.nchunk * 3 * L.fold_npad;
  L.fold_row0[0] = 0; L.fold_row0[1] = (int)pass_rows;
  const int64_t rows = impl == 0 ? pass_rows * L.fold_passes : (int64_t)9 * L.nchunk * L.npad;
  std::vector<uint16_t> wp((size_t)rows * 64, 0);
  if (impl == 0) {
    for (int ps = 0; ps < L.fold_passes; ++ps) pack_layer_fold(L, ps, weight_oihw, wp.data());
  } else {
    pack_layer(L, weight_oihw, wp.data());
  }
  std::vector<float> bpad(64, 0.f);
  std::copy(bias, bias + cout, bpad.begin());
  Batch b;
  TileGeom t{};
  t.lv[0].h = H; t.lv[0].w = W; t.lv[0].pitch = W + 1;
  t.lv[1] = t.lv[0]; t.lv[2] = t.lv[0];
  b.tiles.push_back(t);
  layout_level(b, 0);
  build_fold_schedule(b, 0, h->num_sms);
  const LevelPlan& lp = b.lv[0];
  const int64_t P = lp.pixels;
  std::vector<uint16_t> xs((size_t)L.nchunk * P * 64, 0);      // [plane][pixel][64]
  const LevelGeom g = b.tiles[0].lv[0];
  for (int c = 0; c < cin; ++c)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        xs[((size_t)(c >> 6) * P + (size_t)g.base + (size_t)y * g.pitch + x) * 64 + (c & 63)] = to16(x_nchw[((size_t)c * H + y) * W + x], fmt);
  uint16_t *d_w = nullptr, *d_x = nullptr;
  float *d_b = nullptr, *d_y = nullptr;
  TileGeom* d_t = nullptr; BlockRef* d_blk = nullptr; FoldBand* d_bands = nullptr; FoldSeg* d_segs = nullptr; int32_t* d_off = nullptr;
  int rc = NESR_OK;
  auto cleanup = [&]() {
    cudaFree(d_w); cudaFree(d_x); cudaFree(d_b); cudaFree(d_y); cudaFree(d_t); cudaFree(d_blk); cudaFree(d_bands); cudaFree(d_segs); cudaFree(d_off);
  };
#define DBG_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(h, NESR_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } } while (0)
  DBG_TRY(cudaMalloc(&d_w, wp.size() * 2));
  DBG_TRY(cudaMalloc(&d_x, xs.size() * 2));
  DBG_TRY(cudaMalloc(&d_b, 64 * 4));
  DBG_TRY(cudaMalloc(&d_y, (size_t)P * 64 * 4));
  DBG_TRY(cudaMalloc(&d_t, sizeof(TileGeom)));
  DBG_TRY(cudaMalloc(&d_blk, lp.blocks.size() * sizeof(BlockRef)));
  DBG_TRY(cudaMalloc(&d_bands, lp.bands.size() * sizeof(FoldBand)));
  DBG_TRY(cudaMalloc(&d_off, lp.cta_off.size() * sizeof(int32_t)));
  DBG_TRY(cudaMalloc(&d_segs, lp.segs.size() * sizeof(FoldSeg)));
  DBG_TRY(cudaMemcpy(d_segs, lp.segs.data(), lp.segs.size() * sizeof(FoldSeg), cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_w, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_x, xs.data(), xs.size() * 2, cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_b, bpad.data(), 64 * 4, cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemset(d_y, 0, (size_t)P * 64 * 4));
  DBG_TRY(cudaMemcpy(d_t, b.tiles.data(), sizeof(TileGeom), cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_blk, lp.blocks.data(), lp.blocks.size() * sizeof(BlockRef), cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_bands, lp.bands.data(), lp.bands.size() * sizeof(FoldBand), cudaMemcpyHostToDevice));
  DBG_TRY(cudaMemcpy(d_off, lp.cta_off.data(), lp.cta_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  ConvParams p{};
  p.blocks = d_blk; p.tiles = d_t; p.nblk = (int)lp.blocks.size(); p.level = 0;
  p.src = d_x; p.src_plane_px = (int)P; p.cin = L.cin16; p.wpack = d_w; p.w_row0 = 0; p.npad = L.npad; p.fmt = fmt;
  p.idesc = umma_idesc_f16(hw_fmt(fmt), (uint32_t)L.npad);
  p.bias = d_b; p.cout = cout; p.lrelu = lrelu; p.dst32a = d_y;
  cudaError_t e = cudaSuccess;
  if (impl == 1) {
    e = launch_conv3x3_simt(p, h->stream);
  } else {
    CUtensorMap am, am8, wm;
    if ((rc = make_map(h, &am, d_x, 64, (int64_t)L.nchunk * P, 136)) || (rc = make_map(h, &am8, d_x, 64, (int64_t)L.nchunk * P, 8)) ||
        (rc = make_map(h, &wm, d_w, 64, rows, 3 * L.fold_npad))) { cleanup(); return rc; }
    p.bands = d_bands; p.segs = d_segs; p.cta_band_off = d_off;
    for (int ps = 0; ps < L.fold_passes && e == cudaSuccess; ++ps) {
      p.npad = L.fold_npad; p.c_off = ps * L.fold_npad; p.bias = d_b + p.c_off;
      p.cout = std::min(cout - p.c_off, L.fold_npad); p.w_row0 = L.fold_row0[ps];
      p.idesc = umma_idesc_f16(hw_fmt(fmt), (uint32_t)L.fold_npad);
      if (p.cout > 0) e = launch_conv3x3_fold(am, am8, wm, p, lp.fold_grid, h->stream);
    }
  }
  h->stats.kernel_launches++;
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cleanup(); return fail(h, NESR_E_CUDA, "debug_conv kernel: %s", cudaGetErrorString(e)); }
  std::vector<float> ys((size_t)P * 64);
  DBG_TRY(cudaMemcpy(ys.data(), d_y, ys.size() * 4, cudaMemcpyDeviceToHost));
  for (int c = 0; c < cout; ++c)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        y_nchw[((size_t)c * H + y) * W + x] = ys[trunk_offset((int)(g.base + (int64_t)y * g.pitch + x), c)];
  cleanup();
#undef DBG_TRY
  return NESR_OK;
}

}  // extern "C"
