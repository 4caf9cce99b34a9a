// epilogue.cuh -- the fused conv epilogue, shared by the tcgen05 kernel and the SIMT validation
// kernel so that both write results through identical code.
//
//   v = acc + bias[n]                          (all convs)
//   v = v >= 0 ? v : 0.2 v                     (LeakyReLU 0.2: RDB conv1-4, conv_up1/2, conv_hr)
//   v = v*s1 + res1[px][n]                     (RDB conv5: 0.2*x5 + x ; conv_body: + feat)
//   v = v*s2 + res2[px][n]                     (third RDB of an RRDB: 0.2*out + rrdb_in)
// and stores v as
//   fp32 trunk copies (dst32a/b), a 16-bit copy at a CHANNEL OFFSET of the destination pixel
//   (torch.cat realised as addressing), optionally replicated 2x2 into the next level's layout
//   (nearest x2 upsample realised as store addressing), or -- last layer -- as clamped,
//   round-half-even BGR u8 pasted at the tile's place in the output frame (halo crop + stitch).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "layout.h"

namespace nesr {

__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt_fp16) {
  if (fmt_fp16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float load16(const void* base, size_t idx, int fmt_fp16) {
  const uint16_t raw = reinterpret_cast<const uint16_t*>(base)[idx];
  if (fmt_fp16) return __half2float(__ushort_as_half(raw));
  return __uint_as_float(static_cast<uint32_t>(raw) << 16);
}

// Geometry of one lane's pixel.
struct PixelRef {
  int P;        // flat pixel index at the layer's level
  int y, x;     // position inside the tile
  bool valid;   // false on pad columns / beyond the tile: nothing is stored
};

__device__ __forceinline__ PixelRef locate(const LevelGeom& g, int P) {
  PixelRef r;
  r.P = P;
  const int pl = P - g.base;
  r.y = pl / g.pitch;
  r.x = pl - r.y * g.pitch;
  r.valid = (pl >= 0) && (r.y < g.h) && (r.x < g.w);
  return r;
}

// Channels [n0, n0+16) of one pixel.
__device__ __forceinline__ void epilogue16(const ConvParams& p, const TileGeom& tg, const PixelRef& px, int n0,
                                           float (&v)[16]) {
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b = __ldg(b4 + i);
    v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
  }
  if (p.lrelu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.f ? v[i] : 0.2f * v[i];
  }
  const size_t t64 = static_cast<size_t>(px.P) * kFeat + p.c_off + n0;
  if (p.res1) {
    const float4* r4 = reinterpret_cast<const float4*>(p.res1 + t64);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 r = r4[i];
      v[4 * i + 0] = fmaf(v[4 * i + 0], p.s1, r.x); v[4 * i + 1] = fmaf(v[4 * i + 1], p.s1, r.y);
      v[4 * i + 2] = fmaf(v[4 * i + 2], p.s1, r.z); v[4 * i + 3] = fmaf(v[4 * i + 3], p.s1, r.w);
    }
  }
  if (p.res2) {
    const float4* r4 = reinterpret_cast<const float4*>(p.res2 + t64);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 r = r4[i];
      v[4 * i + 0] = fmaf(v[4 * i + 0], p.s2, r.x); v[4 * i + 1] = fmaf(v[4 * i + 1], p.s2, r.y);
      v[4 * i + 2] = fmaf(v[4 * i + 2], p.s2, r.z); v[4 * i + 3] = fmaf(v[4 * i + 3], p.s2, r.w);
    }
  }
  if (p.dst32a) {
    float4* d4 = reinterpret_cast<float4*>(p.dst32a + t64);
#pragma unroll
    for (int i = 0; i < 4; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (p.dst32b) {
    float4* d4 = reinterpret_cast<float4*>(p.dst32b + t64);
#pragma unroll
    for (int i = 0; i < 4; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (p.dst16) {
    uint4 lo, hi;
    lo.x = pack2(v[0], v[1], p.dst16_fmt);   lo.y = pack2(v[2], v[3], p.dst16_fmt);
    lo.z = pack2(v[4], v[5], p.dst16_fmt);   lo.w = pack2(v[6], v[7], p.dst16_fmt);
    hi.x = pack2(v[8], v[9], p.dst16_fmt);   hi.y = pack2(v[10], v[11], p.dst16_fmt);
    hi.z = pack2(v[12], v[13], p.dst16_fmt); hi.w = pack2(v[14], v[15], p.dst16_fmt);
    uint16_t* base = reinterpret_cast<uint16_t*>(p.dst16) + p.dst16_coff + n0;
    if (!p.dst16_up) {
      uint4* d = reinterpret_cast<uint4*>(base + static_cast<size_t>(px.P) * p.dst16_pitch);
      d[0] = lo; d[1] = hi;
    } else {
      const LevelGeom g2 = tg.lv[p.level + 1];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const size_t P2 = static_cast<size_t>(g2.base) + static_cast<size_t>(2 * px.y + a) * g2.pitch + (2 * px.x + b);
          uint4* d = reinterpret_cast<uint4*>(base + P2 * p.dst16_pitch);
          d[0] = lo; d[1] = hi;
        }
    }
  }
  if (n0 == 0 && (p.out_u8 || p.out_f32)) {
    const int cy = px.y - tg.crop_y, cx = px.x - tg.crop_x;
    if (cy >= 0 && cy < tg.crop_h && cx >= 0 && cx < tg.crop_w) {
      const int Y = tg.out_y0 + cy, X = tg.out_x0 + cx;
      if (p.out_u8) {
        uint8_t* o = p.out_u8 + tg.frame * p.out_frame_stride + Y * p.out_stride + static_cast<int64_t>(X) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < p.cout) {
            const float cl = fminf(fmaxf(v[c], 0.f), 1.f);
            o[2 - c] = static_cast<uint8_t>(rintf(__fmul_rn(cl, 255.f)));     // RGB -> BGR, half-to-even
          }
        }
      } else {
        const size_t plane = static_cast<size_t>(p.out_h) * p.out_w;
        float* o = p.out_f32 + static_cast<size_t>(tg.frame) * p.cout * plane + static_cast<size_t>(Y) * p.out_w + X;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < p.cout) o[c * plane] = v[c];
      }
    }
  }
}

}  // namespace nesr
