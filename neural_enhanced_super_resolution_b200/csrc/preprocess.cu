// Pre-process stage on the GPU: non-local-means denoise in Lab + CLAHE on L, bit-exact with cv2 4.13.
//
// Replaces SuperResolutionPipeline._preprocess_image (reference nesr/nesr.py:668-689):
//
//   if denoise_level > 0: image = cv2.fastNlMeansDenoisingColored(image, None, h, hColor, 7, 21)
//   lab = cvtColor(image, RGB2LAB); L = createCLAHE(2.0, (8, 8)).apply(L); image = cvtColor(lab, LAB2RGB)
//
// Everything here is integer or explicitly rounded float32 arithmetic restated from cv2 (oracle/preprocess.py is the
// specification, pinned against the reference method and against cv2 on all 2^24 colours):
//
//   to_lab_kernel        RGB u8 -> L plane + interleaved (a, b) plane   (gamma table, 3x3 <<12 matrix, cube-root table)
//   nlm_kernel<C>        21x21 search / 7x7 template non-local means on a 1- or 2-channel u8 plane: per thread a column
//                        of 16 output rows, the 7x7 squared-difference sum slides down the column (7 new differences per
//                        row), weights from a shared-memory table indexed by dist >> 6
//   relab_kernel         Lab -> linear BGR -> (sRGB) RGB2LAB in one pass (the hand-over from the denoiser to CLAHE)
//   clahe_lut_kernel     one block per grid tile: histogram (BORDER_REFLECT_101 padding to a grid multiple), clip,
//                        redistribute, cumulative LUT
//   clahe_apply_kernel   bilinear blend of the four surrounding tile LUTs (float32, no FMA contraction), then LAB2RGB
//
// All kernels are HBM/shared-memory stencils (bandwidth / instruction bound, not tensor-pipe work).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "kernels.h"

namespace nesr {
namespace {

#include "lab_tables.inc"

__device__ uint16_t d_gamma_srgb[256];
__device__ uint16_t d_lab_cbrt[3072];
__device__ uint16_t d_lab_to_yf[512];
__device__ uint8_t d_inv_gamma_srgb[4096];

constexpr int kLabShift = 12, kLabShift2 = 15, kBase = 1 << 14, kInvShift = 14;
constexpr int kLScale = 296, kLShift = -1336934;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// cvtColor {RGB,BGR,LRGB,LBGR}2Lab for u8 (cv2 RGB2Lab_b): c[blue_idx] is blue.
__device__ __forceinline__ void rgb_to_lab(int c0, int c1, int c2, int blue_idx, bool srgb, int& L, int& a, int& b) {
  const int r8 = blue_idx == 0 ? c2 : c0, g8 = c1, b8 = blue_idx == 0 ? c0 : c2;
  const int R = srgb ? d_gamma_srgb[r8] : r8 * 8, G = srgb ? d_gamma_srgb[g8] : g8 * 8, B = srgb ? d_gamma_srgb[b8] : b8 * 8;
  const int fX = d_lab_cbrt[descale(R * 1777 + G * 1541 + B * 778, kLabShift)];
  const int fY = d_lab_cbrt[descale(R * 871 + G * 2929 + B * 296, kLabShift)];
  const int fZ = d_lab_cbrt[descale(R * 73 + G * 448 + B * 3575, kLabShift)];
  L = clamp_u8(descale(kLScale * fY + kLShift, kLabShift2));
  a = clamp_u8(descale(500 * (fX - fY) + 128 * (1 << kLabShift2), kLabShift2));
  b = clamp_u8(descale(200 * (fY - fZ) + 128 * (1 << kLabShift2), kLabShift2));
}

// cv2 abToXZ_b evaluated directly (C integer division truncates toward zero, as there).
__device__ __forceinline__ int ab_to_xz(int i) {
  if (i <= 3390) return i * 108 / 841 - kBase * 16 / 116 * 108 / 841;
  return (i * i / kBase) * i / kBase;
}

// cvtColor Lab2{RGB,BGR,LRGB,LBGR} for u8 (cv2 Lab2RGBinteger).
__device__ __forceinline__ void lab_to_rgb(int L, int a, int b, int blue_idx, bool srgb, int& c0, int& c1, int& c2) {
  const int y = d_lab_to_yf[L * 2], ify = d_lab_to_yf[L * 2 + 1];
  const int adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * kBase / 500;
  const int bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * kBase / 200 + 1;
  const long long x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
  auto channel = [&](int k0, int k1, int k2) {
    long long v = (k0 * x + k1 * (long long)y + k2 * z + (1 << (kInvShift - 1))) >> kInvShift;
    const int vi = v < 0 ? 0 : (v > 4095 ? 4095 : (int)v);
    return srgb ? (int)d_inv_gamma_srgb[vi] : (vi * 255) >> 12;
  };
  const int r = channel(12615, -6296, -2223), g = channel(-3773, 7684, 185), bl = channel(217, -836, 4715);
  c0 = blue_idx == 0 ? bl : r; c1 = g; c2 = blue_idx == 0 ? r : bl;
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - i : i;
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise colour kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void to_lab_kernel(const uint8_t* __restrict__ src, int64_t n_px, int blue_idx, int srgb, uint8_t* __restrict__ Lp,
                              uint8_t* __restrict__ abp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_px; i += (int64_t)gridDim.x * blockDim.x) {
    int L, a, b;
    rgb_to_lab(src[3 * i], src[3 * i + 1], src[3 * i + 2], blue_idx, srgb != 0, L, a, b);
    Lp[i] = (uint8_t)L;
    abp[2 * i] = (uint8_t)a; abp[2 * i + 1] = (uint8_t)b;
  }
}

// denoised Lab -> Lab2LBGR (channel 0 blue, linear) -> RGB2LAB (channel 2 blue, sRGB): the image the reference hands to CLAHE
__global__ void relab_kernel(const uint8_t* __restrict__ Lin, const uint8_t* __restrict__ abin, int64_t n_px, uint8_t* __restrict__ Lp,
                             uint8_t* __restrict__ abp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_px; i += (int64_t)gridDim.x * blockDim.x) {
    int c0, c1, c2, L, a, b;
    lab_to_rgb(Lin[i], abin[2 * i], abin[2 * i + 1], 0, false, c0, c1, c2);
    rgb_to_lab(c0, c1, c2, 2, true, L, a, b);
    Lp[i] = (uint8_t)L;
    abp[2 * i] = (uint8_t)a; abp[2 * i + 1] = (uint8_t)b;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// non-local means (cv2 FastNlMeansDenoisingInvoker, template 7, search 21)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kNlmT = 3, kNlmS = 10, kNlmB = kNlmT + kNlmS;        // half template, half search, border
constexpr int kNlmTX = 32, kNlmRows = 16, kNlmTY = 4;               // block: 32 columns x (4 threads x 16 rows)
constexpr int kNlmTileW = kNlmTX + 2 * kNlmB, kNlmTileH = kNlmTY * kNlmRows + 2 * kNlmB;
constexpr int kNlmMaxW = 2048;                                       // weight-table entries kept in shared memory

template <int C>
__global__ void __launch_bounds__(kNlmTX* kNlmTY) nlm_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                             const int* __restrict__ wtab, int n_w) {
  constexpr int kPitch = kNlmTileW * C;
  __shared__ uint8_t tile[kNlmTileH * kPitch];
  __shared__ int s_w[kNlmMaxW];
  const int bx0 = blockIdx.x * kNlmTX, by0 = blockIdx.y * kNlmTY * kNlmRows;
  const int tid = threadIdx.y * kNlmTX + threadIdx.x;
  for (int i = tid; i < kNlmTileH * kNlmTileW; i += kNlmTX * kNlmTY) {
    const int ty = i / kNlmTileW, tx = i - ty * kNlmTileW;
    const int sy = reflect101(by0 + ty - kNlmB, H), sx = reflect101(bx0 + tx - kNlmB, W);
    const uint8_t* p = src + ((int64_t)sy * W + sx) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) tile[ty * kPitch + tx * C + c] = p[c];
  }
  const int n_s = n_w < kNlmMaxW ? n_w : kNlmMaxW;
  for (int i = tid; i < n_s; i += kNlmTX * kNlmTY) s_w[i] = wtab[i];
  __syncthreads();

  const int x = threadIdx.x, y0 = threadIdx.y * kNlmRows;            // tile-relative output column / first row (without border)
  int est[kNlmRows][C];
  int wsum[kNlmRows];
#pragma unroll
  for (int j = 0; j < kNlmRows; ++j) {
    wsum[j] = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) est[j][c] = 0;
  }
  // squared differences of one template row: A = row r around x, B = row r + dy around x + dx (tile coordinates incl. border)
  auto row_dist = [&](const uint8_t* pa, const uint8_t* pb) {
    int s = 0;
#pragma unroll
    for (int t = 0; t < (2 * kNlmT + 1) * C; ++t) {
      const int d = (int)pa[t] - (int)pb[t];
      s += d * d;
    }
    return s;
  };
  for (int dy = -kNlmS; dy <= kNlmS; ++dy) {
    for (int dx = -kNlmS; dx <= kNlmS; ++dx) {
      const uint8_t* pa = tile + (y0 + kNlmB - kNlmT) * kPitch + (x + kNlmB - kNlmT) * C;
      const uint8_t* pb = pa + dy * kPitch + dx * C;
      const uint8_t* pc = tile + (y0 + kNlmB + dy) * kPitch + (x + kNlmB + dx) * C;      // the pixel that gets averaged
      int ring[2 * kNlmT + 1];
      int dist = 0;
#pragma unroll
      for (int r = 0; r < 2 * kNlmT + 1; ++r) {
        ring[r] = row_dist(pa + r * kPitch, pb + r * kPitch);
        dist += ring[r];
      }
#pragma unroll
      for (int j = 0; j < kNlmRows; ++j) {
        if (j > 0) {
          const int fresh = row_dist(pa + (j + 2 * kNlmT) * kPitch, pb + (j + 2 * kNlmT) * kPitch);
          dist += fresh - ring[(j - 1) % (2 * kNlmT + 1)];
          ring[(j - 1) % (2 * kNlmT + 1)] = fresh;
        }
        const int ad = dist >> 6;                                    // 49 -> 64: cv2's almost_template_window_size_sq_bin_shift
        if (ad < n_w) {
          const int w = ad < kNlmMaxW ? s_w[ad] : __ldg(wtab + ad);
          wsum[j] += w;
#pragma unroll
          for (int c = 0; c < C; ++c) est[j][c] += w * (int)pc[j * kPitch + c];
        }
      }
    }
  }
  const int gx = bx0 + x;
  if (gx < W) {
#pragma unroll
    for (int j = 0; j < kNlmRows; ++j) {
      const int gy = by0 + y0 + j;
      if (gy < H) {
        const unsigned ws = (unsigned)wsum[j];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const unsigned v = ((unsigned)est[j][c] + ws / 2) / ws;
          dst[((int64_t)gy * W + gx) * C + c] = (uint8_t)(v > 255u ? 255u : v);
        }
      }
    }
  }
}


// Packed variant (the default): a thread owns FOUR adjacent columns x R rows.  One template row of the four columns needs the
// absolute differences of 10 adjacent bytes: three __vabsdiffu4 on words assembled with funnel shifts (the byte alignment of
// the shifted row depends only on dx, so it is warp-uniform), and the four 7-wide sums of squares come from eight dp4a on
// byte-masked copies.  Channels are kept as separate byte planes in shared memory (a, b de-interleaved at load).
// ~23 instructions per output pixel and search offset instead of ~50, 2 shared-memory loads instead of 14.
constexpr int kNlm4RowsL = 16, kNlm4RowsAB = 8;                       // rows per thread (register budget: 4 x R x (1 + P) accumulators)
constexpr int kN4TX = 32, kN4TY = 4, kN4Cols = 4 * kN4TX;           // block: 128 columns x (4 threads x R rows)
constexpr int kN4TileW = kN4Cols + 2 * kNlmB;                        // 154 bytes used per row
constexpr int kN4PitchW = (kN4TileW + 3) / 4 + 1;                    // 40 words: the fourth word of the last B load stays inside

template <int P, int R>
__global__ void __launch_bounds__(kN4TX* kN4TY) nlm4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                            const int* __restrict__ wtab, int n_w) {
  constexpr int kTH = kN4TY * R + 2 * kNlmB;
  __shared__ uint32_t tile[P][kTH * kN4PitchW];
  __shared__ int s_w[kNlmMaxW];
  const int bx0 = blockIdx.x * kN4Cols, by0 = blockIdx.y * kN4TY * R;
  const int tid = threadIdx.y * kN4TX + threadIdx.x;
  for (int i = tid; i < kTH * kN4PitchW * 4; i += kN4TX * kN4TY) {
    const int ty = i / (kN4PitchW * 4), tx = i - ty * (kN4PitchW * 4);
    const int sy = reflect101(by0 + ty - kNlmB, H), sx = reflect101(bx0 + tx - kNlmB, W);
    const uint8_t* p = src + ((int64_t)sy * W + sx) * P;
#pragma unroll
    for (int c = 0; c < P; ++c) reinterpret_cast<uint8_t*>(tile[c])[i] = p[c];
  }
  for (int i = tid; i < kNlmMaxW; i += kN4TX * kN4TY) s_w[i] = i < n_w ? wtab[i] : 0;
  __syncthreads();

  const int y0 = threadIdx.y * R;
  int est[R][4][P];
  int wsum[R][4];
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      wsum[j][c] = 0;
#pragma unroll
      for (int p = 0; p < P; ++p) est[j][c][p] = 0;
    }
  for (int dy = -kNlmS; dy <= kNlmS; ++dy) {
    for (int dx = -kNlmS; dx <= kNlmS; ++dx) {
      const int sb = ((kNlmS + dx) & 3) * 8, wb = threadIdx.x + ((kNlmS + dx) >> 2);          // shifted row: first byte = column + 10 + dx
      const int sc = ((kNlmB + dx) & 3) * 8, wc = threadIdx.x + ((kNlmB + dx) >> 2);          // averaged pixel: column + 13 + dx
      // sums of squared differences of tile row `ra` (vs row ra + dy) over the 7 template columns of each of the 4 outputs
      auto row4 = [&](int ra, unsigned (&s)[4]) {
        s[0] = s[1] = s[2] = s[3] = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const uint32_t* ta = tile[p] + ra * kN4PitchW + threadIdx.x + 2;
          const uint32_t* tb = tile[p] + (ra + dy) * kN4PitchW + wb;
          const uint32_t a_0 = ta[0], a_1 = ta[1], a_2 = ta[2];
          const uint32_t b_0 = tb[0], b_1 = tb[1], b_2 = tb[2], b_3 = tb[3];
          const uint32_t d0 = __vabsdiffu4(__funnelshift_r(a_0, a_1, 16), __funnelshift_r(b_0, b_1, sb));
          const uint32_t d1 = __vabsdiffu4(__funnelshift_r(a_1, a_2, 16), __funnelshift_r(b_1, b_2, sb));
          const uint32_t d2 = __vabsdiffu4(a_2 >> 16, __funnelshift_r(b_2, b_3, sb)) & 0x0000ffffu;
          const uint32_t q1 = __dp4a(d1, d1, 0u);
          const uint32_t m1 = d1 & 0x00ffffffu, m01 = d0 & 0xffffff00u, m02 = d0 & 0xffff0000u, m03 = d0 & 0xff000000u, e8 = d2 & 0xffu;
          s[0] += __dp4a(d0, d0, __dp4a(m1, m1, 0u));
          s[1] += __dp4a(m01, m01, q1);
          s[2] += __dp4a(m02, m02, __dp4a(e8, e8, q1));
          s[3] += __dp4a(m03, m03, __dp4a(d2, d2, q1));
        }
      };
      unsigned ring[2 * kNlmT + 1][4];
      unsigned dist[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int r = 0; r < 2 * kNlmT + 1; ++r) {
        row4(y0 + kNlmB - kNlmT + r, ring[r]);
#pragma unroll
        for (int c = 0; c < 4; ++c) dist[c] += ring[r][c];
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (j > 0) {
          unsigned fresh[4];
          row4(y0 + kNlmB + kNlmT + j, fresh);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            dist[c] += fresh[c] - ring[(j - 1) % (2 * kNlmT + 1)][c];
            ring[(j - 1) % (2 * kNlmT + 1)][c] = fresh[c];
          }
        }
        int w[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const unsigned ad = dist[c] >> 6;                         // 49 -> 64: cv2's almost_template_window_size_sq_bin_shift
          w[c] = s_w[ad < (unsigned)kNlmMaxW ? ad : kNlmMaxW - 1];   // the table's tail is zero (n_w < kNlmMaxW, checked by the launcher)
        }
        if ((w[0] | w[1] | w[2] | w[3]) != 0) {
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const uint32_t* tc = tile[p] + (y0 + j + kNlmB + dy) * kN4PitchW + wc;
            const uint32_t cw = __funnelshift_r(tc[0], tc[1], sc);
#pragma unroll
            for (int c = 0; c < 4; ++c) est[j][c][p] += w[c] * (int)((cw >> (8 * c)) & 0xffu);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) wsum[j][c] += w[c];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int gy = by0 + y0 + j;
    if (gy >= H) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gx = bx0 + threadIdx.x * 4 + c;
      if (gx >= W) continue;
      const unsigned ws = (unsigned)wsum[j][c];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const unsigned v = ((unsigned)est[j][c][p] + ws / 2) / ws;
        dst[((int64_t)gy * W + gx) * P + p] = (uint8_t)(v > 255u ? 255u : v);
      }
    }
  }
}

template <int P, int R>
void launch_nlm4(const uint8_t* src, uint8_t* dst, int H, int W, const int* wtab, int n_w, cudaStream_t stream) {
  const dim3 grid((W + kN4Cols - 1) / kN4Cols, (H + kN4TY * R - 1) / (kN4TY * R)), block(kN4TX, kN4TY);
  nlm4_kernel<P, R><<<grid, block, 0, stream>>>(src, dst, H, W, wtab, n_w);
}

// ------------------------------------------------------------------------------------------------------------------
// CLAHE (cv2 clahe.cpp)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kClaheThreads = 1024;                                 // four 256-thread groups, one private histogram each

__global__ void __launch_bounds__(kClaheThreads) clahe_lut_kernel(const uint8_t* __restrict__ L, int H, int W, int tiles_x, int tw, int th,
                                                                  int clip_limit, float lut_scale, uint8_t* __restrict__ luts) {
  __shared__ int part[kClaheThreads / 256][256];
  __shared__ int hist[256];
  __shared__ int s_clipped;
  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int t = threadIdx.x, grp = t >> 8;
  part[grp][t & 255] = 0;
  if (t == 0) s_clipped = 0;
  __syncthreads();
  for (int i = t; i < tw * th; i += kClaheThreads) {
    const int yy = i / tw, xx = i - yy * tw;
    const int sy = tile_y * th + yy, sx = tile_x * tw + xx;         // coordinates in the padded plane
    const int py = sy < H ? sy : reflect101(sy, H), px = sx < W ? sx : reflect101(sx, W);
    atomicAdd(&part[grp][L[(int64_t)py * W + px]], 1);
  }
  __syncthreads();
  const bool bin = t < 256;                                          // one thread per histogram bin from here on
  if (bin) {
    int v = 0;
#pragma unroll
    for (int g = 0; g < kClaheThreads / 256; ++g) v += part[g][t];
    hist[t] = v;
  }
  __syncthreads();
  if (clip_limit > 0) {
    if (bin) {
      const int hv = hist[t];
      if (hv > clip_limit) {
        atomicAdd(&s_clipped, hv - clip_limit);
        hist[t] = clip_limit;
      }
    }
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    const int residual = clipped - batch * 256;
    if (bin) {
      int v = hist[t] + batch;
      if (residual != 0) {
        const int step = 256 / residual > 1 ? 256 / residual : 1;
        // entries 0, step, 2*step, ... get one more, `residual` of them at most
        if (t % step == 0 && t / step < residual) v += 1;
      }
      hist[t] = v;
    }
    __syncthreads();
  }
  if (t == 0) {
    int sum = 0;
    uint8_t* lut = luts + (int64_t)blockIdx.x * 256;
    for (int i = 0; i < 256; ++i) {
      sum += hist[i];
      lut[i] = (uint8_t)clamp_u8(__float2int_rn(__fmul_rn((float)sum, lut_scale)));
    }
  }
}

__global__ void clahe_apply_kernel(const uint8_t* __restrict__ Lp, const uint8_t* __restrict__ abp, int H, int W, int tiles_x, int tiles_y,
                                   float inv_tw, float inv_th, const uint8_t* __restrict__ luts, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
  int tx1 = (int)floorf(txf);
  const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
  int tx2 = tx1 + 1 < tiles_x - 1 ? tx1 + 1 : tiles_x - 1;
  tx1 = tx1 > 0 ? tx1 : 0;
  const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int ty1 = (int)floorf(tyf);
  const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
  int ty2 = ty1 + 1 < tiles_y - 1 ? ty1 + 1 : tiles_y - 1;
  ty1 = ty1 > 0 ? ty1 : 0;
  const int64_t i = (int64_t)y * W + x;
  const int v = Lp[i];
  const float l11 = luts[(ty1 * tiles_x + tx1) * 256 + v], l12 = luts[(ty1 * tiles_x + tx2) * 256 + v];
  const float l21 = luts[(ty2 * tiles_x + tx1) * 256 + v], l22 = luts[(ty2 * tiles_x + tx2) * 256 + v];
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
  const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  const int Lnew = clamp_u8(__float2int_rn(res));
  int c0, c1, c2;
  lab_to_rgb(Lnew, abp[2 * i], abp[2 * i + 1], 2, true, c0, c1, c2);
  out[3 * i] = (uint8_t)c0; out[3 * i + 1] = (uint8_t)c1; out[3 * i + 2] = (uint8_t)c2;
}

bool g_tables_uploaded[64] = {};
std::mutex g_tables_mutex;                                     // engines on several threads may race to the first upload

// The consuming kernels run on non-blocking streams, which have no implicit ordering with the legacy stream these pageable
// copies use (and a pageable copy may return before its DMA has landed): the device is synchronised before the flag is set.
cudaError_t upload_tables() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  if (dev < 64 && g_tables_uploaded[dev]) return cudaSuccess;
  if ((e = cudaMemcpyToSymbol(d_gamma_srgb, kGammaSrgb, sizeof(kGammaSrgb))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(d_lab_cbrt, kLabCbrt, sizeof(kLabCbrt))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(d_lab_to_yf, kLabToYF, sizeof(kLabToYF))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(d_inv_gamma_srgb, kInvGammaSrgb, sizeof(kInvGammaSrgb))) != cudaSuccess) return e;
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
  if (dev < 64) g_tables_uploaded[dev] = true;
  return cudaSuccess;
}

}  // namespace

// cv2's almost_dist2weight_ (non-zero prefix): weights for dist >> 6, DistSquared::calcWeight.
std::vector<int32_t> nlm_weight_table(float h, int channels) {
  const int fixed_point_mult = 2147483647 / (21 * 21 * 255);
  const double mult = 64.0 / 49.0;
  const int n = (int)(255.0 * 255.0 * channels / mult + 1);
  const float den = h * h * (float)channels;
  std::vector<int32_t> tab;
  for (int ad = 0; ad < n; ++ad) {
    const double dist = ad * mult;
    double w = std::exp(-dist / den);
    if (std::isnan(w)) w = 1.0;
    int weight = (int)std::nearbyint(fixed_point_mult * w);
    if (weight < 0.001 * fixed_point_mult) weight = 0;
    if (weight == 0 && ad > 0) break;
    tab.push_back(weight);
  }
  return tab;
}

const void* lab_table_host(int which, int* count, int* elem_bytes) {
  switch (which) {
    case 0: *count = 256; *elem_bytes = 2; return kGammaSrgb;
    case 1: *count = 3072; *elem_bytes = 2; return kLabCbrt;
    case 2: *count = 512; *elem_bytes = 2; return kLabToYF;
    case 3: *count = 4096; *elem_bytes = 1; return kInvGammaSrgb;
    default: *count = 0; *elem_bytes = 0; return nullptr;
  }
}

size_t preprocess_workspace_bytes(int H, int W, int tiles_x, int tiles_y) {
  const size_t n = ((size_t)H * W + 255) & ~(size_t)255;
  return 6 * n + (size_t)tiles_x * tiles_y * 256 + 256;
}

// rgb (device, H x W x 3) -> out (device, H x W x 3).  wtab_l / wtab_ab: device weight tables (non-zero prefixes), nullptr = no denoise.
cudaError_t launch_preprocess(const uint8_t* rgb, uint8_t* out, int H, int W, const int32_t* wtab_l, int n_wl, const int32_t* wtab_ab, int n_wab,
                              float clip, int tiles_x, int tiles_y, uint8_t* workspace, int* launches, cudaStream_t stream) {
  cudaError_t e = upload_tables();
  if (e != cudaSuccess) return e;
  const int64_t n_px = (int64_t)H * W;
  const size_t n = ((size_t)n_px + 255) & ~(size_t)255;
  uint8_t* L0 = workspace;
  uint8_t* ab0 = L0 + n;
  uint8_t* L1 = ab0 + 2 * n;
  uint8_t* ab1 = L1 + n;
  uint8_t* luts = ab1 + 2 * n;
  const int ew_blocks = (int)((n_px + 255) / 256 < 148 * 16 ? (n_px + 255) / 256 : 148 * 16);
  int nl = 0;
  const bool denoise = wtab_l != nullptr && wtab_ab != nullptr;
  if (denoise) {
    to_lab_kernel<<<ew_blocks, 256, 0, stream>>>(rgb, n_px, 0, 0, L0, ab0);                 // LBGR2Lab: channel 0 read as blue, linear
    const dim3 grid((W + kNlmTX - 1) / kNlmTX, (H + kNlmTY * kNlmRows - 1) / (kNlmTY * kNlmRows)), block(kNlmTX, kNlmTY);
    static const int impl = getenv("NESR_B200_NLM_IMPL") ? atoi(getenv("NESR_B200_NLM_IMPL")) : 0;   // 1: the per-column kernel
    if (impl == 1 || n_wl >= kNlmMaxW) nlm_kernel<1><<<grid, block, 0, stream>>>(L0, L1, H, W, wtab_l, n_wl);
    else if (impl == 2) launch_nlm4<1, 8>(L0, L1, H, W, wtab_l, n_wl, stream);
    else launch_nlm4<1, kNlm4RowsL>(L0, L1, H, W, wtab_l, n_wl, stream);
    if (impl == 1 || n_wab >= kNlmMaxW) nlm_kernel<2><<<grid, block, 0, stream>>>(ab0, ab1, H, W, wtab_ab, n_wab);
    else launch_nlm4<2, kNlm4RowsAB>(ab0, ab1, H, W, wtab_ab, n_wab, stream);
    relab_kernel<<<ew_blocks, 256, 0, stream>>>(L1, ab1, n_px, L0, ab0);
    nl += 4;
  } else {
    to_lab_kernel<<<ew_blocks, 256, 0, stream>>>(rgb, n_px, 2, 1, L0, ab0);                 // RGB2LAB
    nl += 1;
  }
  const int ext_w = W % tiles_x == 0 && H % tiles_y == 0 ? W : W + tiles_x - W % tiles_x;
  const int ext_h = W % tiles_x == 0 && H % tiles_y == 0 ? H : H + tiles_y - H % tiles_y;
  const int tw = ext_w / tiles_x, th = ext_h / tiles_y;
  const int area = tw * th;
  int clip_limit = 0;
  if (clip > 0.0f) {
    clip_limit = (int)((double)clip * area / 256);
    if (clip_limit < 1) clip_limit = 1;
  }
  const float lut_scale = 255.0f / (float)area;
  clahe_lut_kernel<<<tiles_x * tiles_y, kClaheThreads, 0, stream>>>(L0, H, W, tiles_x, tw, th, clip_limit, lut_scale, luts);
  const dim3 ablock(32, 8), agrid((W + 31) / 32, (H + 7) / 8);
  clahe_apply_kernel<<<agrid, ablock, 0, stream>>>(L0, ab0, H, W, tiles_x, tiles_y, 1.0f / (float)tw, 1.0f / (float)th, luts, out);
  nl += 2;
  if (launches) *launches = nl;
  return cudaGetLastError();
}

}  // namespace nesr
