// stencil.cu -- the two integer post-process kernels, bit-exact with the reference's numpy / cv2
// arithmetic (restated in oracle/postprocess.py):
//
//   blend   : SuperResolutionPipeline._ensemble_results  (reference nesr/nesr.py:1047-1054)
//             acc_f32 = f32( f64(acc_f32) + f64(img) * w_i )  per member, then truncate to u8.
//   sharpen : SuperResolutionPipeline._postprocess_image  (reference nesr/nesr.py:1062-1080)
//             gray = (9798 R + 19235 G + 3735 B + 2^14) >> 15
//             G2 = 13-tap, G3 = 19-tap separable Q8.8 Gaussians, BORDER_REFLECT_101, one rounding
//             mask = max(gray - G2(gray), 0) > 10
//             out = mask ? sat_u8(round_half_even(1.5*img - 0.5*G3(img))) : img
// The sharpen kernel is a single shared-memory-tiled pass: each image byte is read from HBM once
// (plus halo) and written once.
#include <cstdlib>

#include "kernels.h"

namespace nesr {

namespace {

// ---------------------------------------------------------------------------------------------
// blend
// ---------------------------------------------------------------------------------------------
// acc_f32 = f32( f64(acc_f32) + f64(byte) * w ) per member, in member order (oracle/postprocess.py ensemble_results)
__device__ __forceinline__ float blend_step(float acc, uint32_t byte, double w) {
  return static_cast<float>(__dadd_rn(static_cast<double>(acc), __dmul_rn(static_cast<double>(byte), w)));
}

// K members known at compile time: pointers and the 16-byte vectors live in registers.  (With K a run-time value the
// per-member arrays were indexed dynamically and lived in local memory: 0.7 TB/s of algorithmic traffic on a B200.)
// The kernel was bound by its conversions (u8 -> f64, f32 -> f64, f64 -> f32: nine per byte for K = 3, on the 16-lane XU pipe), so the
// products come from shared-memory tables built once per block with the same instruction the per-byte path used:
//   s_prod[m][b] = __dmul_rn(f64(b), w[m])                     (the reference's  img.astype(f32) * weights[m], a float64 product)
//   s_first[b]   = f32(f64(0.f) + s_prod[0][b])                (the first member's step: 0 + P is exact, so this is f32(P))
// and a later member's step is  acc = f32(f64(acc) + s_prod[m][b])  -- the same roundings in the same order, bit for bit.
// Two independent 16-byte vectors per member are in flight per thread and iteration.
template <int K>
__global__ void __launch_bounds__(256) blend_kernel_k(const BlendParams p, const int vec_ok) {
  __shared__ double s_prod[K][256];
  __shared__ float s_first[256];
  for (int i = threadIdx.x; i < 256 * K; i += 256) s_prod[i >> 8][i & 255] = __dmul_rn(static_cast<double>(i & 255), p.weights[i >> 8]);
  __syncthreads();
  s_first[threadIdx.x] = static_cast<float>(__dadd_rn(0.0, s_prod[0][threadIdx.x]));
  __syncthreads();
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = vec_ok ? p.nbytes / 16 : 0;
  const uint4* src[K];
#pragma unroll
  for (int m = 0; m < K; ++m) src[m] = reinterpret_cast<const uint4*>(p.members[m]);
  uint4* dst = reinterpret_cast<uint4*>(p.out);
  auto blend1 = [&](const uint32_t (&b)[K]) {                   // one output byte from the K member bytes
    float acc = s_first[b[0]];
#pragma unroll
    for (int m = 1; m < K; ++m) acc = static_cast<float>(__dadd_rn(static_cast<double>(acc), s_prod[m][b[m]]));
    return static_cast<uint32_t>(static_cast<uint8_t>(acc));    // astype(uint8): truncation (acc >= 0)
  };
  auto blend16 = [&](const uint4 (&in)[K]) {
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      uint32_t word = 0;
#pragma unroll
      for (int by = 0; by < 4; ++by) {
        uint32_t b[K];
#pragma unroll
        for (int m = 0; m < K; ++m) b[m] = (reinterpret_cast<const uint32_t*>(&in[m])[wd] >> (8 * by)) & 0xFF;
        word |= blend1(b) << (8 * by);
      }
      ow[wd] = word;
    }
    return o;
  };
  int64_t i = tid;
  for (; i + nthreads < nvec; i += 2 * nthreads) {
    uint4 a[K], b[K];
#pragma unroll
    for (int m = 0; m < K; ++m) { a[m] = __ldcs(src[m] + i); b[m] = __ldcs(src[m] + i + nthreads); }
    __stcs(dst + i, blend16(a));
    __stcs(dst + i + nthreads, blend16(b));
  }
  for (; i < nvec; i += nthreads) {
    uint4 a[K];
#pragma unroll
    for (int m = 0; m < K; ++m) a[m] = __ldcs(src[m] + i);
    __stcs(dst + i, blend16(a));
  }
  for (int64_t j = nvec * 16 + tid; j < p.nbytes; j += nthreads) {
    uint32_t b[K];
#pragma unroll
    for (int m = 0; m < K; ++m) b[m] = p.members[m][j];
    p.out[j] = static_cast<uint8_t>(blend1(b));
  }
}

// Two members with weights exactly 0.5 (the reference's 1/K for K = 2): f32(a * 0.5) and f32(acc + b * 0.5) are exact (multiples of
// 0.5 below 256), so the truncated result is floor((a + b) / 2) -- one SIMD halving add per four bytes ("K=2 => exactly
// (a+b)>>1").  No conversions, no f64: the kernel is a pure stream of 16-byte loads and stores, four vectors in flight per thread.
__global__ void __launch_bounds__(256) blend_half_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out,
                                                         const int64_t nbytes, const int vec_ok) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = vec_ok ? nbytes / 16 : 0;
  const uint4* va = reinterpret_cast<const uint4*>(a);
  const uint4* vb = reinterpret_cast<const uint4*>(b);
  uint4* vo = reinterpret_cast<uint4*>(out);
  auto avg = [](const uint4& x, const uint4& y) {
    return make_uint4(__vhaddu4(x.x, y.x), __vhaddu4(x.y, y.y), __vhaddu4(x.z, y.z), __vhaddu4(x.w, y.w));
  };
  int64_t i = tid;
  for (; i + 3 * nthreads < nvec; i += 4 * nthreads) {
    uint4 x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { x[k] = __ldcs(va + i + k * nthreads); y[k] = __ldcs(vb + i + k * nthreads); }
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(vo + i + k * nthreads, avg(x[k], y[k]));
  }
  for (; i < nvec; i += nthreads) __stcs(vo + i, avg(__ldcs(va + i), __ldcs(vb + i)));
  for (int64_t j = nvec * 16 + tid; j < nbytes; j += nthreads) out[j] = static_cast<uint8_t>((a[j] + b[j]) >> 1);
}

// any K up to kMaxBlendMembers (members indexed at run time)
__global__ void __launch_bounds__(256) blend_kernel(const BlendParams p, const int vec_ok) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = vec_ok ? p.nbytes / 16 : 0;
  for (int64_t i = tid; i < nvec; i += nthreads) {
    float acc[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) acc[b] = 0.f;
    for (int m = 0; m < p.k; ++m) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.members[m]) + i);
      const double w = p.weights[m];
      const uint32_t* vw = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
      for (int b = 0; b < 16; ++b) acc[b] = blend_step(acc[b], (vw[b >> 2] >> (8 * (b & 3))) & 0xFF, w);
    }
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      uint32_t word = 0;
#pragma unroll
      for (int by = 0; by < 4; ++by) word |= static_cast<uint32_t>(static_cast<uint8_t>(acc[wd * 4 + by])) << (8 * by);
      ow[wd] = word;
    }
    reinterpret_cast<uint4*>(p.out)[i] = o;
  }
  for (int64_t i = nvec * 16 + tid; i < p.nbytes; i += nthreads) {
    float acc = 0.f;
    for (int m = 0; m < p.k; ++m) acc = blend_step(acc, p.members[m][i], p.weights[m]);
    p.out[i] = static_cast<uint8_t>(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// sharpen
// ---------------------------------------------------------------------------------------------
constexpr int kTW = 64, kTH = 32;         // output tile
constexpr int kR3 = 9, kR2 = 6;           // radii of the 19- and 13-tap kernels
constexpr int kInH = kTH + 2 * kR3;       // 50
constexpr int kInWp = 88;                 // padded pitch in bytes (22 aligned words)
constexpr int kGrayRows = kTH + 2 * kR2;  // 44
constexpr int kRuns = kTW / 4;            // runs of four horizontally adjacent outputs per tile row
constexpr int kTp = 52;                   // pitch of the transposed row sums: 13 words (odd: conflict-free across columns)

// Both passes run on dp4a.  Four outputs o0 .. o0+3 (o0 a multiple of 4) read the 24 bytes o0 .. o0+23 of a line as six ALIGNED
// words; output j weighs byte i with q[i - j] (zero outside the kernel), so each output is six dp4a with weight words that depend
// only on (j, word) -- no byte extraction, no unaligned loads.  All weights are < 128 and non-negative.
//   rows   : lines are tile rows of the planar input bytes; the 16-bit sums are stored as two BYTE planes (high, low), transposed
//            ([column][row]), so that
//   columns: lines are tile columns of those byte planes: sum = 256 * dp4a(high bytes) + dp4a(low bytes), exact.
struct TapWeights {
  uint32_t w3[4][6];   // 19-tap kernel (taps 0 and 18 are zero)
  uint32_t w2r[4][6];  // 13-tap kernel inside the 19-tap window of a ROW (offset kR3 - kR2 = 3)
  uint32_t w2c[4][4];  // 13-tap kernel on the 44 gray row sums of a COLUMN (no offset: 16-byte window)
};
constexpr TapWeights make_tap_weights() {
  constexpr int q3[19] = {0, 1, 3, 4, 9, 14, 20, 28, 32, 34, 32, 28, 20, 14, 9, 4, 3, 1, 0};
  constexpr int q2[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
  TapWeights r{};
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 6; ++k) {
      uint32_t a = 0, b = 0, c = 0;
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * k + e;
        const int t3 = i - j, t2 = i - j - (kR3 - kR2), tc = i - j;
        if (t3 >= 0 && t3 < 19) a |= static_cast<uint32_t>(q3[t3]) << (8 * e);
        if (t2 >= 0 && t2 < 13) b |= static_cast<uint32_t>(q2[t2]) << (8 * e);
        if (tc >= 0 && tc < 13) c |= static_cast<uint32_t>(q2[tc]) << (8 * e);
      }
      r.w3[j][k] = a; r.w2r[j][k] = b;
      if (k < 4) r.w2c[j][k] = c;
    }
  return r;
}
__constant__ TapWeights c_tw = make_tap_weights();

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// One shared-memory tiled pass: every image byte is read from HBM once (plus halo) and written once.  The arithmetic is the
// integer restatement of cv2's fixed-point Gaussians (oracle/postprocess.py); sums are exact, so their order is free:
//   load   : tile + 9-pixel halo -> planar R, G, B and gray bytes (BORDER_REFLECT_101)
//   rows   : 19-tap (three channels) and 13-tap (gray) horizontal sums, Q8.8 in 16 bits
//   columns: vertical sums, one rounding, threshold mask, unsharp with round-half-even, saturate
//
// kExtMask (the segmentation-masked unsharp of the reference, nesr/nesr.py:728-747): the same unsharp -- GaussianBlur(img, sigma 3),
// addWeighted(img, 1.5, blurred, -0.5) -- applied where cv2.dilate(object_mask, ones(3, 3)) == 1 instead of where the image has
// detail: `ext_mask` is the H x W u8 object mask at image resolution, its 3 x 3 dilation (a maximum over the pixels INSIDE the
// image: cv2's default border value for dilate never wins) is taken here, and the gray plane / 13-tap blur are not computed.
template <bool kExtMask>
__global__ void __launch_bounds__(256) sharpen_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                      const int H, const int W, const int bgr, const uint8_t* __restrict__ ext_mask) {
  __shared__ __align__(16) uint8_t s_in[4][kInH][kInWp];     // planar R, G, B, gray
  __shared__ __align__(16) uint8_t s_t3[3][2][kTW][kTp];     // row sums of the 19-tap blur: [channel][high | low byte][column][row]
  __shared__ __align__(16) uint8_t s_t2[2][kTW][kTp];        // row sums of the 13-tap blur on gray: [high | low][column][row - 3]

  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const int tid = threadIdx.x;

  // load: a thread takes four horizontally adjacent pixels of a tile row (one index computation, one 32-bit store per plane;
  // columns 82, 83 of the 84 loaded are never weighted)
  for (int idx = tid; idx < kInH * (kInWp / 4 - 1); idx += 256) {
    const int ly = idx / (kInWp / 4 - 1), lx = 4 * (idx - ly * (kInWp / 4 - 1));
    const int gy = reflect101(y0 + ly - kR3, H);
    const uint8_t* row = in + static_cast<size_t>(gy) * W * 3;
    const int gx0 = x0 + lx - kR3;
    uint32_t pr = 0, pg = 0, pb = 0, py = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gx = (gx0 >= 0 && gx0 + 3 < W) ? gx0 + e : reflect101(gx0 + e, W);
      const uint8_t* px = row + gx * 3;
      const int c0 = px[0], c1 = px[1], c2 = px[2];
      const int r = bgr ? c2 : c0, b = bgr ? c0 : c2;
      pr |= static_cast<uint32_t>(r) << (8 * e); pg |= static_cast<uint32_t>(c1) << (8 * e); pb |= static_cast<uint32_t>(b) << (8 * e);
      if (!kExtMask) py |= static_cast<uint32_t>((9798 * r + 19235 * c1 + 3735 * b + 16384) >> 15) << (8 * e);
    }
    *reinterpret_cast<uint32_t*>(&s_in[0][ly][lx]) = pr;
    *reinterpret_cast<uint32_t*>(&s_in[1][ly][lx]) = pg;
    *reinterpret_cast<uint32_t*>(&s_in[2][ly][lx]) = pb;
    if (!kExtMask) *reinterpret_cast<uint32_t*>(&s_in[3][ly][lx]) = py;
  }
  __syncthreads();

  // rows: item = (plane, run of four outputs, tile row); consecutive threads take consecutive rows, so the transposed byte
  // stores of a warp fall into consecutive bytes
  for (int idx = tid; idx < (kExtMask ? 3 : 4) * kRuns * kInH; idx += 256) {
    const int plane = idx / (kRuns * kInH);
    const int rem = idx - plane * (kRuns * kInH);
    const int run = rem / kInH, ly = rem - run * kInH;
    if (plane == 3 && (ly < kR3 - kR2 || ly >= kR3 - kR2 + kGrayRows)) continue;
    const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(&s_in[plane][ly][4 * run]);
    uint32_t wv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) wv[k] = wsrc[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t a = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) a = __dp4a(wv[k], plane == 3 ? c_tw.w2r[j][k] : c_tw.w3[j][k], a);
      if (plane == 3) {
        s_t2[0][4 * run + j][ly - (kR3 - kR2)] = static_cast<uint8_t>(a >> 8);
        s_t2[1][4 * run + j][ly - (kR3 - kR2)] = static_cast<uint8_t>(a);
      } else {
        s_t3[plane][0][4 * run + j][ly] = static_cast<uint8_t>(a >> 8);
        s_t3[plane][1][4 * run + j][ly] = static_cast<uint8_t>(a);
      }
    }
  }
  __syncthreads();

  // columns: item = (group of four vertically adjacent outputs, column); consecutive threads take consecutive columns
  for (int idx = tid; idx < (kTH / 4) * kTW; idx += 256) {
    const int og = idx / kTW, ox = idx - og * kTW;
    const int oy0 = 4 * og;
    const int gx = x0 + ox;
    if (gx >= W || y0 + oy0 >= H) continue;
    bool mask[4];
    bool any = false;
    if (kExtMask) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gy = y0 + oy0 + j;
        int m = 0;
        if (gy < H)
          for (int yy = max(gy - 1, 0); yy <= min(gy + 1, H - 1); ++yy)
            for (int xx = max(gx - 1, 0); xx <= min(gx + 1, W - 1); ++xx) m = max(m, static_cast<int>(__ldg(ext_mask + static_cast<size_t>(yy) * W + xx)));
        mask[j] = m == 1;                                      // np.where(mask == 1, ...): a label other than 1 selects nothing
        any |= mask[j];
      }
    } else {
      const uint32_t* hi = reinterpret_cast<const uint32_t*>(&s_t2[0][ox][oy0]);
      const uint32_t* lo = reinterpret_cast<const uint32_t*>(&s_t2[1][ox][oy0]);
      uint32_t hv[4], lv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { hv[k] = hi[k]; lv[k] = lo[k]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t sh = 0, sl = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { sh = __dp4a(hv[k], c_tw.w2c[j][k], sh); sl = __dp4a(lv[k], c_tw.w2c[j][k], sl); }
        const int g2 = static_cast<int>((sh * 256u + sl + 32768u) >> 16);
        const int gray = s_in[3][oy0 + j + kR3][ox + kR3];
        mask[j] = (gray - g2) > 10;                            // saturating subtract then threshold
        any |= mask[j];
      }
    }
    int res[4][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int b3[4] = {0, 0, 0, 0};
      if (any) {
        const uint32_t* hi = reinterpret_cast<const uint32_t*>(&s_t3[c][0][ox][oy0]);
        const uint32_t* lo = reinterpret_cast<const uint32_t*>(&s_t3[c][1][ox][oy0]);
        uint32_t hv[6], lv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { hv[k] = hi[k]; lv[k] = lo[k]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t sh = 0, sl = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) { sh = __dp4a(hv[k], c_tw.w3[j][k], sh); sl = __dp4a(lv[k], c_tw.w3[j][k], sl); }
          b3[j] = static_cast<int>((sh * 256u + sl + 32768u) >> 16);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = s_in[c][oy0 + j + kR3][ox + kR3];
        int v = a;
        if (mask[j]) {
          const int t2 = 3 * a - b3[j];                        // twice (1.5 a - 0.5 b)
          const int half = t2 >> 1;                            // floor
          v = half + ((t2 & 1) & (half & 1));                  // ties to even
          v = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        res[j][c] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gy = y0 + oy0 + j;
      if (gy < H) {
        uint8_t* o = out + (static_cast<size_t>(gy) * W + gx) * 3;
        o[0] = static_cast<uint8_t>(bgr ? res[j][2] : res[j][0]);
        o[1] = static_cast<uint8_t>(res[j][1]);
        o[2] = static_cast<uint8_t>(bgr ? res[j][0] : res[j][2]);
      }
    }
  }
}

}  // namespace

cudaError_t launch_blend(const BlendParams& p, cudaStream_t stream) {
  if (p.nbytes <= 0) return cudaSuccess;
  int vec_ok = (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  for (int m = 0; m < p.k; ++m) vec_ok &= (reinterpret_cast<uintptr_t>(p.members[m]) & 15) == 0;
  const int64_t work = vec_ok ? (p.nbytes + 15) / 16 : p.nbytes;
  int64_t blocks = (work + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const unsigned g = static_cast<unsigned>(blocks);
  if (p.k == 2 && p.weights[0] == 0.5 && p.weights[1] == 0.5) {
    blend_half_kernel<<<g, 256, 0, stream>>>(p.members[0], p.members[1], p.out, p.nbytes, vec_ok);
    return cudaGetLastError();
  }
  switch (p.k) {
    case 1: blend_kernel_k<1><<<g, 256, 0, stream>>>(p, vec_ok); break;
    case 2: blend_kernel_k<2><<<g, 256, 0, stream>>>(p, vec_ok); break;
    case 3: blend_kernel_k<3><<<g, 256, 0, stream>>>(p, vec_ok); break;
    case 4: blend_kernel_k<4><<<g, 256, 0, stream>>>(p, vec_ok); break;
    default: blend_kernel<<<g, 256, 0, stream>>>(p, vec_ok); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_sharpen(const uint8_t* in, uint8_t* out, int32_t H, int32_t W, int32_t bgr, const uint8_t* ext_mask, cudaStream_t stream) {
  if (H <= 0 || W <= 0) return cudaSuccess;
  const char* impl = getenv("NESR_B200_SHARPEN_IMPL");          // 1: the dp4a kernel below (cross-check); default: sharpen_mma.cu
  if (!impl || atoi(impl) != 1) return launch_sharpen_mma(in, out, H, W, bgr, ext_mask, stream);
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH);
  if (ext_mask) sharpen_kernel<true><<<grid, 256, 0, stream>>>(in, out, H, W, bgr, ext_mask);
  else sharpen_kernel<false><<<grid, 256, 0, stream>>>(in, out, H, W, bgr, nullptr);
  return cudaGetLastError();
}

}  // namespace nesr
