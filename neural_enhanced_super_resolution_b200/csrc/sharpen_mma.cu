// sharpen_mma.cu -- the adaptive-sharpen post-process (and its segmentation-masked variant) with both Gaussian passes on the
// TENSOR pipe.  Same arithmetic as stencil.cu's dp4a kernel (the integer restatement of cv2's fixed-point GaussianBlur in
// oracle/postprocess.py; reference nesr/nesr.py:1056-1084 and :732-747): sums of products of bytes are exact in s32, so their
// order is free and the result is bit-identical.
//
// A separable Q8.8 Gaussian pass is a product with a banded Toeplitz matrix.  The 19-tap kernel's outer taps are zero, so 16
// outputs read 16 + 16 = 32 inputs: one mma.sync.m16n8k32 (u8 x u8 -> s32; SASS IMMA.16832.U8.U8) per 16 outputs x 8 lines,
// with the A operand -- A[m][k] = q[k - m] -- a CONSTANT held in four registers per thread:
//   rows    : D[out column][row] = A x B,  B[k][n] = planar input byte (row n, column c0 + k): four consecutive bytes of a row
//             are one B register, so the B fragment is two aligned 32-bit shared-memory loads;
//             the 16-bit row sums are stored as two byte planes (high, low), TRANSPOSED ([column][row]), so that
//   columns : D[out row][column] = A x B,  B[k][n] = row-sum byte (column n, row r0 + k): again two aligned loads;
//             sum = 256 * (A x high) + (A x low), one rounding (+ 2^15) >> 16.
// Pitches of 80 B (input rows) and 48 B (transposed columns) make both fragment loads bank-conflict free (pitch in words = 4 mod 8).
// The dp4a kernel issued >= 68 dp4a per pixel and was bound by instruction issue (6 % of the HBM roofline: DESIGN.md 4.4); here
// a 64 x 32 tile is 96 + 128 IMMAs.  The thread that holds a pixel's four blurred values (R, G, B at sigma 3, gray at sigma 2:
// the same accumulator coordinates in all eight column MMAs) evaluates mask and unsharp in registers.
// Global traffic: aligned 32-bit loads of the interleaved bytes (de-interleaved with byte permutes) and 32-bit stores from a
// staged output tile when the row pitch allows it (W % 4 == 0); bytes with BORDER_REFLECT_101 otherwise and on border tiles.
#include "kernels.h"

namespace nesr {

namespace {

constexpr int kTW = 64, kTH = 32;          // output tile
constexpr int kHalo = 8;                   // radius of the 17 non-zero taps of the sigma-3 kernel (sigma 2: 6)
constexpr int kInW = kTW + 2 * kHalo;      // 80 bytes: row pitch of the planar input, 20 words
constexpr int kInH = kTH + 2 * kHalo;      // 48 bytes: column pitch of the transposed row sums, 12 words
constexpr int kOutPitch = kTW * 3 + 16;    // 208 bytes = 52 words (4 mod 8): the eight rows a warp writes at once fall into different banks; 16-byte rows

// The A fragments of the two Toeplitz matrices, per lane (g = lane / 4, t = lane % 4; register r: row g + 8 (r & 1), k = 4t + 16 (r >> 1) + i in
// byte i):  A[m][k] = tap[k - m], zero outside the kernel.  sigma 3: 19 taps of which the 17 inner ones are non-zero (tap index k - m);
// sigma 2: 13 taps centred in the same 17-tap window (tap index k - m - 2).  Built at compile time and read with ONE 16-byte load per
// lane and matrix (indexed __constant__ bytes cost 64 divergent constant loads per thread: 60 % of the kernel's stall samples).
struct FragTable { uint32_t a[2][32][4]; };
constexpr FragTable make_frags() {
  constexpr int t3[17] = {1, 3, 4, 9, 14, 20, 28, 32, 34, 32, 28, 20, 14, 9, 4, 3, 1};
  constexpr int t2[13] = {1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1};
  FragTable f{};
  for (int lane = 0; lane < 32; ++lane)
    for (int r = 0; r < 4; ++r) {
      const int m = lane / 4 + 8 * (r & 1), k0 = 4 * (lane % 4) + 16 * (r >> 1);
      uint32_t w3 = 0, w2 = 0;
      for (int i = 0; i < 4; ++i) {
        const int d = k0 + i - m;
        if (d >= 0 && d < 17) w3 |= static_cast<uint32_t>(t3[d]) << (8 * i);
        if (d >= 2 && d < 15) w2 |= static_cast<uint32_t>(t2[d - 2]) << (8 * i);
      }
      f.a[0][lane][r] = w3; f.a[1][lane][r] = w2;
    }
  return f;
}
__device__ const FragTable g_frags = make_frags();

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// D (16 x 8, s32) += A (16 x 32, u8, row) x B (32 x 8, u8, col).  Fragments (PTX ISA, m16n8k32 8-bit): g = lane / 4, t = lane % 4;
//   a0: row g, k 4t..4t+3 | a1: row g+8, same k | a2: row g, k 16+4t.. | a3: row g+8, k 16+4t..
//   b0: k 4t..4t+3, column g | b1: k 16+4t.., column g
//   d0: (row g, column 2t) d1: (g, 2t+1) d2: (g+8, 2t) d3: (g+8, 2t+1)
__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool kExtMask>
__global__ void __launch_bounds__(256) sharpen_mma_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int H, const int W,
                                                          const int bgr, const uint8_t* __restrict__ ext_mask, const int aligned4, const int aligned16) {
  __shared__ __align__(16) uint8_t s_in[4][kInH][kInW];          // planar channel 0, 1, 2 (memory order) and gray
  __shared__ __align__(16) uint8_t s_t[4][2][kTW][kInH];         // row sums: [plane][high | low byte][column][row]
  __shared__ __align__(16) uint8_t s_out[kTH][kOutPitch];        // interleaved output tile

  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wr = bgr ? 3735 : 9798, wb = bgr ? 9798 : 3735;      // gray weights of memory channels 0 and 2

  // ---- load: item = four horizontally adjacent pixels of a tile row -> one word per plane
  const bool fast_x = aligned4 && x0 - kHalo >= 0 && x0 + kTW + kHalo <= W;
  if (fast_x) {
    // interior tile, 4-byte aligned rows: three aligned words hold the four pixels (c0 c1 c2 c0 | c1 c2 c0 c1 | c2 c0 c1 c2); the planes
    // come out of byte permutes and gray out of dp4a on the interleaved words: 9798 = 38 * 256 + 70, 19235 = 75 * 256 + 35,
    // 3735 = 14 * 256 + 151, so gray = (256 * dp4a(px, high weights) + dp4a(px, low weights) + 2^14) >> 15, exactly.
    const uint32_t h0 = static_cast<uint32_t>(wr) >> 8, h1 = 75u, h2 = static_cast<uint32_t>(wb) >> 8;
    const uint32_t l0 = static_cast<uint32_t>(wr) & 255u, l1 = 35u, l2 = static_cast<uint32_t>(wb) & 255u;
    const uint32_t hA = h0 | h1 << 8 | h2 << 16, hB0 = h0 << 24, hB1 = h1 | h2 << 8, hC1 = h0 << 16 | h1 << 24, hC2 = h2, hD = h0 << 8 | h1 << 16 | h2 << 24;
    const uint32_t lA = l0 | l1 << 8 | l2 << 16, lB0 = l0 << 24, lB1 = l1 | l2 << 8, lC1 = l0 << 16 | l1 << 24, lC2 = l2, lD = l0 << 8 | l1 << 16 | l2 << 24;
    // 240 threads take 12 tile rows x 20 items per step, four steps: a thread keeps its column and walks down 12 rows at a time, so every
    // address is a base plus a step-constant offset; all twelve loads of a thread are issued before the first is used
    constexpr int kRowsPerStep = 12, kSteps = kInH / kRowsPerStep;
    static_assert(kRowsPerStep * (kInW / 4) <= 256 && kSteps * kRowsPerStep == kInH, "load mapping");
    if (tid < kRowsPerStep * (kInW / 4)) {
      const int r = tid / (kInW / 4), lx = 4 * (tid - r * (kInW / 4));
      const bool fast_y = y0 - kHalo >= 0 && y0 + kTH + kHalo <= H;
      const uint8_t* col = in + static_cast<size_t>(x0 + lx - kHalo) * 3;
      uint32_t w[kSteps][3];
#pragma unroll
      for (int it = 0; it < kSteps; ++it) {
        const int ty = y0 + r + kRowsPerStep * it - kHalo;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(col + static_cast<size_t>(fast_y ? ty : reflect101(ty, H)) * W * 3);
        w[it][0] = __ldg(src); w[it][1] = __ldg(src + 1); w[it][2] = __ldg(src + 2);
      }
      uint8_t* dst = &s_in[0][r][lx];
#pragma unroll
      for (int it = 0; it < kSteps; ++it) {
        const uint32_t w0 = w[it][0], w1 = w[it][1], w2 = w[it][2];
        uint8_t* d = dst + it * kRowsPerStep * kInW;
        *reinterpret_cast<uint32_t*>(d) = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);                       // bytes 0, 3, 6, 9 of the 12
        *reinterpret_cast<uint32_t*>(d + kInH * kInW) = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);         // bytes 1, 4, 7, 10
        *reinterpret_cast<uint32_t*>(d + 2 * kInH * kInW) = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);     // bytes 2, 5, 8, 11
        if (!kExtMask) {
          const uint32_t g0 = (__dp4a(w0, hA, 0u) * 256u + __dp4a(w0, lA, 16384u)) >> 15;
          const uint32_t g1 = (__dp4a(w1, hB1, __dp4a(w0, hB0, 0u)) * 256u + __dp4a(w1, lB1, __dp4a(w0, lB0, 16384u))) >> 15;
          const uint32_t g2 = (__dp4a(w2, hC2, __dp4a(w1, hC1, 0u)) * 256u + __dp4a(w2, lC2, __dp4a(w1, lC1, 16384u))) >> 15;
          const uint32_t g3 = (__dp4a(w2, hD, 0u) * 256u + __dp4a(w2, lD, 16384u)) >> 15;
          *reinterpret_cast<uint32_t*>(d + 3 * kInH * kInW) = g0 | g1 << 8 | g2 << 16 | g3 << 24;
        }
      }
    }
  } else {
    for (int idx = tid; idx < kInH * (kInW / 4); idx += 256) {
      const int ly = idx / (kInW / 4), lx = 4 * (idx - ly * (kInW / 4));
      const int gy = reflect101(y0 + ly - kHalo, H);
      const uint8_t* row = in + static_cast<size_t>(gy) * W * 3;
      const int gx0 = x0 + lx - kHalo;
      uint32_t p0 = 0, p1 = 0, p2 = 0, py = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint8_t* px = row + static_cast<size_t>(reflect101(gx0 + e, W)) * 3;
        const int c0 = px[0], c1 = px[1], c2 = px[2];
        p0 |= static_cast<uint32_t>(c0) << (8 * e); p1 |= static_cast<uint32_t>(c1) << (8 * e); p2 |= static_cast<uint32_t>(c2) << (8 * e);
        py |= static_cast<uint32_t>((wr * c0 + 19235 * c1 + wb * c2 + 16384) >> 15) << (8 * e);
      }
      *reinterpret_cast<uint32_t*>(&s_in[0][ly][lx]) = p0;
      *reinterpret_cast<uint32_t*>(&s_in[1][ly][lx]) = p1;
      *reinterpret_cast<uint32_t*>(&s_in[2][ly][lx]) = p2;
      if (!kExtMask) *reinterpret_cast<uint32_t*>(&s_in[3][ly][lx]) = py;
    }
  }
  uint32_t a3[4], a2[4];
  {
    const uint4 f3 = __ldg(reinterpret_cast<const uint4*>(g_frags.a[0][lane])), f2 = __ldg(reinterpret_cast<const uint4*>(g_frags.a[1][lane]));
    a3[0] = f3.x; a3[1] = f3.y; a3[2] = f3.z; a3[3] = f3.w;
    a2[0] = f2.x; a2[1] = f2.y; a2[2] = f2.z; a2[3] = f2.w;
  }
  __syncthreads();

  // ---- rows: unit = (plane, 16 output columns, 8 tile rows); a warp keeps its 16 columns and three of the six row groups, so that every
  // address below is one base plus a compile-time offset
  {
    const int mb = warp & 3, nb0 = (warp >> 2) * 3;
    const uint8_t* src0 = &s_in[0][8 * nb0 + g][16 * mb + 4 * t];
    uint8_t* dst0 = &s_t[0][0][16 * mb + g][8 * nb0 + 2 * t];
#pragma unroll
    for (int plane = 0; plane < (kExtMask ? 3 : 4); ++plane)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(src0 + plane * (kInH * kInW) + j * 8 * kInW);
        int d[4] = {0, 0, 0, 0};
        if (plane == 3) imma(d, a2, src[0], src[4]); else imma(d, a3, src[0], src[4]);
        // d0, d1: column 16 mb + g, rows 8 nb + 2t, + 1;  d2, d3: column + 8
        uint8_t* dst = dst0 + plane * (2 * kTW * kInH) + j * 8;
        *reinterpret_cast<uint16_t*>(dst) = static_cast<uint16_t>(__byte_perm(d[0], d[1], 0x0051));
        *reinterpret_cast<uint16_t*>(dst + kTW * kInH) = static_cast<uint16_t>(__byte_perm(d[0], d[1], 0x0040));
        *reinterpret_cast<uint16_t*>(dst + 8 * kInH) = static_cast<uint16_t>(__byte_perm(d[2], d[3], 0x0051));
        *reinterpret_cast<uint16_t*>(dst + kTW * kInH + 8 * kInH) = static_cast<uint16_t>(__byte_perm(d[2], d[3], 0x0040));
      }
  }
  __syncthreads();

  // ---- columns + mask + unsharp: block = (16 output rows, 8 columns); the thread's four pixels are rows oy, oy + 8, columns ox, ox + 1
  static_assert(kTW / 8 == 8, "one 8-column block per warp");
#pragma unroll
  for (int vb = 0; vb < kTH / 16; ++vb) {
    const int cb = warp;
    const int ox = 8 * cb + 2 * t;
    auto column_blur = [&](int plane, const uint32_t (&af)[4], int (&out4)[4]) {      // the block's four blurred values of one plane
      const uint32_t* hi = reinterpret_cast<const uint32_t*>(&s_t[plane][0][8 * cb + g][16 * vb + 4 * t]);
      const uint32_t* lo = reinterpret_cast<const uint32_t*>(&s_t[plane][1][8 * cb + g][16 * vb + 4 * t]);
      int dh[4] = {0, 0, 0, 0}, dl[4] = {32768, 32768, 32768, 32768};   // the rounding constant rides in the accumulator
      imma(dh, af, hi[0], hi[4]);
      imma(dl, af, lo[0], lo[4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) out4[i] = static_cast<int>((static_cast<uint32_t>(dh[i]) * 256u + static_cast<uint32_t>(dl[i])) >> 16);
    };
    // the mask of the thread's four pixels first: where no pixel of the warp's 16 x 8 block is selected (flat regions of a photo, most of
    // a segmentation mask's background) the three colour blurs are not needed and the block is a copy
    bool mask[2][2];
    if (kExtMask) {
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gy = y0 + 16 * vb + g + 8 * hrow, gx = x0 + ox + e;
          int m = 0;
          if (gy < H && gx < W)
            for (int yy = max(gy - 1, 0); yy <= min(gy + 1, H - 1); ++yy)
              for (int xx = max(gx - 1, 0); xx <= min(gx + 1, W - 1); ++xx) m = max(m, static_cast<int>(__ldg(ext_mask + static_cast<size_t>(yy) * W + xx)));
          mask[hrow][e] = m == 1;                                // np.where(mask == 1, ...): another label selects nothing
        }
    } else {
      int g2[4];
      column_blur(3, a2, g2);
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const uint32_t gray2 = *reinterpret_cast<const uint16_t*>(&s_in[3][16 * vb + g + 8 * hrow + kHalo][ox + kHalo]);
        mask[hrow][0] = static_cast<int>(gray2 & 255) - g2[2 * hrow] > 10;             // saturating subtract, then threshold
        mask[hrow][1] = static_cast<int>(gray2 >> 8) - g2[2 * hrow + 1] > 10;
      }
    }
    const bool any = __any_sync(0xffffffffu, mask[0][0] | mask[0][1] | mask[1][0] | mask[1][1]);
    int blur[3][4] = {};
    if (any) {                                                   // warp-uniform: the MMAs below are warp-wide
#pragma unroll
      for (int plane = 0; plane < 3; ++plane) column_blur(plane, a3, blur[plane]);
    }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int oy = 16 * vb + g + 8 * hrow;
      uint32_t res[2][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t a2px = *reinterpret_cast<const uint16_t*>(&s_in[c][oy + kHalo][ox + kHalo]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = static_cast<int>((a2px >> (8 * e)) & 255);
          int v = a;
          if (mask[hrow][e]) {
            const int t2 = 3 * a - blur[c][2 * hrow + e];        // twice (1.5 a - 0.5 b)
            const int half = t2 >> 1;                            // floor
            v = half + ((t2 & 1) & (half & 1));                  // ties to even
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
          }
          res[e][c] = static_cast<uint32_t>(v);
        }
      }
      uint16_t* o = reinterpret_cast<uint16_t*>(&s_out[oy][ox * 3]);           // six bytes: c0 c1 c2 c0 c1 c2
      o[0] = static_cast<uint16_t>(res[0][0] | (res[0][1] << 8));
      o[1] = static_cast<uint16_t>(res[0][2] | (res[1][0] << 8));
      o[2] = static_cast<uint16_t>(res[1][1] | (res[1][2] << 8));
    }
  }
  __syncthreads();

  // ---- store
  const int rows = min(kTH, H - y0), cols = min(kTW, W - x0);
  if (aligned16 && cols == kTW) {                                // 12 x 16 bytes per row
    for (int idx = tid; idx < rows * (kTW * 3 / 16); idx += 256) {
      const int oy = idx / (kTW * 3 / 16), q = idx - oy * (kTW * 3 / 16);
      reinterpret_cast<uint4*>(out + (static_cast<size_t>(y0 + oy) * W + x0) * 3)[q] = reinterpret_cast<const uint4*>(&s_out[oy][0])[q];
    }
  } else if (aligned4 && cols == kTW) {
    for (int idx = tid; idx < rows * (kTW * 3 / 4); idx += 256) {
      const int oy = idx / (kTW * 3 / 4), wd = idx - oy * (kTW * 3 / 4);
      reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(y0 + oy) * W + x0) * 3)[wd] = reinterpret_cast<const uint32_t*>(&s_out[oy][0])[wd];
    }
  } else {
    for (int idx = tid; idx < rows * cols * 3; idx += 256) {
      const int oy = idx / (cols * 3), b = idx - oy * (cols * 3);
      out[(static_cast<size_t>(y0 + oy) * W + x0) * 3 + b] = s_out[oy][b];
    }
  }
}

}  // namespace

cudaError_t launch_sharpen_mma(const uint8_t* in, uint8_t* out, int32_t H, int32_t W, int32_t bgr, const uint8_t* ext_mask, cudaStream_t stream) {
  if (H <= 0 || W <= 0) return cudaSuccess;
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH);
  const int aligned4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3) == 0;
  const int aligned16 = (W % 16 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;      // rows of the output are 16-byte aligned
  if (ext_mask) sharpen_mma_kernel<true><<<grid, 256, 0, stream>>>(in, out, H, W, bgr, ext_mask, aligned4, aligned16);
  else sharpen_mma_kernel<false><<<grid, 256, 0, stream>>>(in, out, H, W, bgr, nullptr, aligned4, aligned16);
  return cudaGetLastError();
}

}  // namespace nesr
