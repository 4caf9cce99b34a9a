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
__device__ __forceinline__ uint8_t blend_byte(const BlendParams& p, const uint8_t (&b)[kMaxBlendMembers]) {
  float acc = 0.f;
  for (int m = 0; m < p.k; ++m)
    acc = static_cast<float>(__dadd_rn(static_cast<double>(acc), __dmul_rn(static_cast<double>(b[m]), p.weights[m])));
  return static_cast<uint8_t>(acc);           // astype(uint8): truncation (acc >= 0)
}

__global__ void __launch_bounds__(256) blend_kernel(const BlendParams p, const int vec_ok) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = vec_ok ? p.nbytes / 16 : 0;
  for (int64_t i = tid; i < nvec; i += nthreads) {
    uint4 in[kMaxBlendMembers];
    for (int m = 0; m < p.k; ++m) in[m] = __ldg(reinterpret_cast<const uint4*>(p.members[m]) + i);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      uint32_t word = 0;
#pragma unroll
      for (int by = 0; by < 4; ++by) {
        uint8_t b[kMaxBlendMembers];
        for (int m = 0; m < p.k; ++m) b[m] = (reinterpret_cast<const uint32_t*>(&in[m])[wd] >> (8 * by)) & 0xFF;
        word |= static_cast<uint32_t>(blend_byte(p, b)) << (8 * by);
      }
      ow[wd] = word;
    }
    reinterpret_cast<uint4*>(p.out)[i] = o;
  }
  for (int64_t i = nvec * 16 + tid; i < p.nbytes; i += nthreads) {
    uint8_t b[kMaxBlendMembers];
    for (int m = 0; m < p.k; ++m) b[m] = p.members[m][i];
    p.out[i] = blend_byte(p, b);
  }
}

// ---------------------------------------------------------------------------------------------
// sharpen
// ---------------------------------------------------------------------------------------------
constexpr int kTW = 64, kTH = 32;         // output tile
constexpr int kR3 = 9, kR2 = 6;           // radii of the 19- and 13-tap kernels
constexpr int kInW = kTW + 2 * kR3;       // 82
constexpr int kInH = kTH + 2 * kR3;       // 50
constexpr int kInWp = kInW + 2;           // padded pitch
constexpr int kGrayRows = kTH + 2 * kR2;  // 44

__constant__ int c_q3[19] = {0, 1, 3, 4, 9, 14, 20, 28, 32, 34, 32, 28, 20, 14, 9, 4, 3, 1, 0};
__constant__ int c_q2[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) sharpen_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                      const int H, const int W, const int bgr) {
  __shared__ uint8_t s_in[3][kInH][kInWp];          // planar R, G, B
  __shared__ uint8_t s_gray[kInH][kInWp];
  __shared__ uint16_t s_h3[3][kInH][kTW];           // row pass of the 19-tap blur (Q8.8, <= 65280)
  __shared__ uint16_t s_h2[kGrayRows][kTW];         // row pass of the 13-tap blur on gray

  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const int tid = threadIdx.x;

  for (int idx = tid; idx < kInH * kInW; idx += 256) {
    const int ly = idx / kInW, lx = idx - ly * kInW;
    const int gy = reflect101(y0 + ly - kR3, H), gx = reflect101(x0 + lx - kR3, W);
    const uint8_t* px = in + (static_cast<size_t>(gy) * W + gx) * 3;
    const int c0 = px[0], c1 = px[1], c2 = px[2];
    const int r = bgr ? c2 : c0, b = bgr ? c0 : c2;
    s_in[0][ly][lx] = r; s_in[1][ly][lx] = c1; s_in[2][ly][lx] = b;
    s_gray[ly][lx] = static_cast<uint8_t>((9798 * r + 19235 * c1 + 3735 * b + 16384) >> 15);
  }
  __syncthreads();

  for (int idx = tid; idx < kInH * kTW; idx += 256) {
    const int ly = idx / kTW, ox = idx - ly * kTW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int acc = 0;
#pragma unroll
      for (int t = 1; t < 18; ++t) acc += c_q3[t] * s_in[c][ly][ox + t];       // taps 0 and 18 are zero
      s_h3[c][ly][ox] = static_cast<uint16_t>(acc);
    }
    if (ly >= kR3 - kR2 && ly < kR3 - kR2 + kGrayRows) {
      int acc = 0;
#pragma unroll
      for (int t = 0; t < 13; ++t) acc += c_q2[t] * s_gray[ly][ox + (kR3 - kR2) + t];
      s_h2[ly - (kR3 - kR2)][ox] = static_cast<uint16_t>(acc);
    }
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
    const int gray = s_gray[oy + kR3][ox + kR3];
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
  blend_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, vec_ok);
  return cudaGetLastError();
}

cudaError_t launch_sharpen(const uint8_t* in, uint8_t* out, int32_t H, int32_t W, int32_t bgr, cudaStream_t stream) {
  if (H <= 0 || W <= 0) return cudaSuccess;
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH);
  sharpen_kernel<<<grid, 256, 0, stream>>>(in, out, H, W, bgr);
  return cudaGetLastError();
}

}  // namespace nesr
