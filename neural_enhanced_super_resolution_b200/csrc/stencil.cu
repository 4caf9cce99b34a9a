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

// K members known at compile time: pointers, weights and the 16-byte vectors live in registers.  (With K a run-time value the
// per-member arrays were indexed dynamically and lived in local memory: 0.7 TB/s of algorithmic traffic on a B200.)
// Two independent 16-byte vectors per member are in flight per thread and iteration.
template <int K>
__global__ void __launch_bounds__(256) blend_kernel_k(const BlendParams p, const int vec_ok) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = vec_ok ? p.nbytes / 16 : 0;
  const uint4* src[K];
  double w[K];
#pragma unroll
  for (int m = 0; m < K; ++m) { src[m] = reinterpret_cast<const uint4*>(p.members[m]); w[m] = p.weights[m]; }
  uint4* dst = reinterpret_cast<uint4*>(p.out);
  auto blend16 = [&](const uint4 (&in)[K]) {
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      uint32_t word = 0;
#pragma unroll
      for (int by = 0; by < 4; ++by) {
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < K; ++m) acc = blend_step(acc, (reinterpret_cast<const uint32_t*>(&in[m])[wd] >> (8 * by)) & 0xFF, w[m]);
        word |= static_cast<uint32_t>(static_cast<uint8_t>(acc)) << (8 * by);   // astype(uint8): truncation (acc >= 0)
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
    float acc = 0.f;
#pragma unroll
    for (int m = 0; m < K; ++m) acc = blend_step(acc, p.members[m][j], w[m]);
    p.out[j] = static_cast<uint8_t>(acc);
  }
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
constexpr int kInW = kTW + 2 * kR3;       // 82
constexpr int kInH = kTH + 2 * kR3;       // 50
constexpr int kInWp = 88;                 // padded pitch in bytes (22 aligned words)
constexpr int kGrayRows = kTH + 2 * kR2;  // 44
constexpr int kRuns = kTW / 4;            // runs of four horizontally adjacent outputs per tile row

__constant__ int c_q3[19] = {0, 1, 3, 4, 9, 14, 20, 28, 32, 34, 32, 28, 20, 14, 9, 4, 3, 1, 0};
__constant__ int c_q2[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};

// Row pass with dp4a.  Four outputs ox0 .. ox0+3 (ox0 a multiple of 4) read the 24 bytes ox0 .. ox0+23 of a tile row as six
// ALIGNED words; output j weighs byte i with q[i - j] (zero outside the kernel), so each output is six dp4a with weight words
// that depend only on (j, word) -- no byte extraction, no unaligned loads.  All weights are < 128 and non-negative.
struct RowWeights {
  uint32_t w3[4][6];   // 19-tap kernel (taps 0 and 18 are zero)
  uint32_t w2[4][6];   // 13-tap kernel, centred in the 19-tap window (offset kR3 - kR2 = 3)
};
constexpr RowWeights make_row_weights() {
  constexpr int q3[19] = {0, 1, 3, 4, 9, 14, 20, 28, 32, 34, 32, 28, 20, 14, 9, 4, 3, 1, 0};
  constexpr int q2[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
  RowWeights r{};
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 6; ++k) {
      uint32_t a = 0, b = 0;
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * k + e;
        const int t3 = i - j, t2 = i - j - (kR3 - kR2);
        if (t3 >= 0 && t3 < 19) a |= static_cast<uint32_t>(q3[t3]) << (8 * e);
        if (t2 >= 0 && t2 < 13) b |= static_cast<uint32_t>(q2[t2]) << (8 * e);
      }
      r.w3[j][k] = a; r.w2[j][k] = b;
    }
  return r;
}
__constant__ RowWeights c_rw = make_row_weights();

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// One shared-memory tiled pass: every image byte is read from HBM once (plus halo) and written once.  The arithmetic is the
// integer restatement of cv2's fixed-point Gaussians (oracle/postprocess.py); sums are exact, so the order is free:
//   load   : tile + 9-pixel halo -> planar R, G, B and gray bytes (BORDER_REFLECT_101)
//   rows   : 19-tap (three channels) and 13-tap (gray) horizontal sums as dp4a over aligned words -> 16-bit Q8.8
//   columns: vertical sums, one rounding, threshold mask, unsharp with round-half-even, saturate
__global__ void __launch_bounds__(256) sharpen_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                      const int H, const int W, const int bgr) {
  __shared__ __align__(16) uint8_t s_in[4][kInH][kInWp];     // planar R, G, B, gray
  __shared__ __align__(16) uint16_t s_h3[3][kInH][kTW];      // row pass of the 19-tap blur (Q8.8, <= 65280)
  __shared__ __align__(16) uint16_t s_h2[kGrayRows][kTW];    // row pass of the 13-tap blur on gray

  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const int tid = threadIdx.x;

  for (int idx = tid; idx < kInH * kInW; idx += 256) {
    const int ly = idx / kInW, lx = idx - ly * kInW;
    const int gy = reflect101(y0 + ly - kR3, H), gx = reflect101(x0 + lx - kR3, W);
    const uint8_t* px = in + (static_cast<size_t>(gy) * W + gx) * 3;
    const int c0 = px[0], c1 = px[1], c2 = px[2];
    const int r = bgr ? c2 : c0, b = bgr ? c0 : c2;
    s_in[0][ly][lx] = r; s_in[1][ly][lx] = c1; s_in[2][ly][lx] = b;
    s_in[3][ly][lx] = static_cast<uint8_t>((9798 * r + 19235 * c1 + 3735 * b + 16384) >> 15);
  }
  __syncthreads();

  // rows: item = (plane, tile row, run of four outputs)
  for (int idx = tid; idx < 4 * kInH * kRuns; idx += 256) {
    const int plane = idx / (kInH * kRuns);
    const int rem = idx - plane * (kInH * kRuns);
    const int ly = rem / kRuns, run = rem - ly * kRuns;
    if (plane == 3 && (ly < kR3 - kR2 || ly >= kR3 - kR2 + kGrayRows)) continue;
    const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(&s_in[plane][ly][4 * run]);
    uint32_t wv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) wv[k] = wsrc[k];
    uint32_t acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t a = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) a = __dp4a(wv[k], plane == 3 ? c_rw.w2[j][k] : c_rw.w3[j][k], a);
      acc[j] = a;
    }
    uint16_t* dst = plane == 3 ? &s_h2[ly - (kR3 - kR2)][4 * run] : &s_h3[plane][ly][4 * run];
    *reinterpret_cast<uint2*>(dst) = make_uint2(acc[0] | (acc[1] << 16), acc[2] | (acc[3] << 16));
  }
  __syncthreads();

  for (int idx = tid; idx < kTH * kTW; idx += 256) {
    const int oy = idx / kTW, ox = idx - oy * kTW;
    const int gy = y0 + oy, gx = x0 + ox;
    if (gy >= H || gx >= W) continue;
    int g2 = 0;
#pragma unroll
    for (int t = 0; t < 13; ++t) g2 += c_q2[t] * s_h2[oy + t][ox];
    g2 = (g2 + 32768) >> 16;
    const int gray = s_in[3][oy + kR3][ox + kR3];
    const bool mask = (gray - g2) > 10;                    // saturating subtract then threshold
    int res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int a = s_in[c][oy + kR3][ox + kR3];
      int v = a;
      if (mask) {
        int b3 = 0;
#pragma unroll
        for (int t = 1; t < 18; ++t) b3 += c_q3[t] * s_h3[c][oy + t][ox];
        b3 = (b3 + 32768) >> 16;
        const int t2 = 3 * a - b3;                         // twice (1.5 a - 0.5 b)
        const int half = t2 >> 1;                          // floor
        v = half + ((t2 & 1) & (half & 1));                // ties to even
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
      }
      res[c] = v;
    }
    uint8_t* o = out + (static_cast<size_t>(gy) * W + gx) * 3;
    o[0] = static_cast<uint8_t>(bgr ? res[2] : res[0]);
    o[1] = static_cast<uint8_t>(res[1]);
    o[2] = static_cast<uint8_t>(bgr ? res[0] : res[2]);
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
  switch (p.k) {
    case 1: blend_kernel_k<1><<<g, 256, 0, stream>>>(p, vec_ok); break;
    case 2: blend_kernel_k<2><<<g, 256, 0, stream>>>(p, vec_ok); break;
    case 3: blend_kernel_k<3><<<g, 256, 0, stream>>>(p, vec_ok); break;
    case 4: blend_kernel_k<4><<<g, 256, 0, stream>>>(p, vec_ok); break;
    default: blend_kernel<<<g, 256, 0, stream>>>(p, vec_ok); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_sharpen(const uint8_t* in, uint8_t* out, int32_t H, int32_t W, int32_t bgr, cudaStream_t stream) {
  if (H <= 0 || W <= 0) return cudaSuccess;
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH);
  sharpen_kernel<<<grid, 256, 0, stream>>>(in, out, H, W, bgr);
  return cudaGetLastError();
}

}  // namespace nesr
