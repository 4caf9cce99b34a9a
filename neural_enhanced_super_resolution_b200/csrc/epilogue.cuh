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

// ---------------------------------------------------------------------------------------------
// fp32 trunk buffers (trunk / rrdb / feat) are touched only by epilogues, one pixel per lane, so they
// use a BLOCKED layout that makes that access coalesced: [pixel/32][channel/8][pixel%32][channel%8].
// A warp reading 8 channels of 32 consecutive pixels moves 1 KB of (nearly) contiguous memory with
// one 256-bit access per lane instead of 32 scattered lines.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t trunk_offset(int P, int ch) {
  return (static_cast<size_t>(P >> 5) * 8 + (ch >> 3)) * 256 + (static_cast<size_t>(P & 31) << 3) + (ch & 7);
}
__device__ __forceinline__ void ldg256(const float* ptr, float* v) {
  uint32_t r[8];
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(ptr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// Streaming variants for the fp32 trunk (touched once per dense block): evict-first in L2, no L1 allocation.
__device__ __forceinline__ void ldg256_stream(const float* ptr, float* v) {
  uint32_t r[8];
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(ptr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void stg256f_stream(float* ptr, const float* v) {
  asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void stg256(void* ptr, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void stg256f(float* ptr, const float* v) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
  stg256(ptr, r);
}
// 16 residual channels [ch0, ch0+16) of pixel P (ch0 multiple of 8).
__device__ __forceinline__ void load_trunk16(const float* buf, int P, int ch0, float* r) {
  ldg256_stream(buf + trunk_offset(P, ch0), r);
  ldg256_stream(buf + trunk_offset(P, ch0 + 8), r + 8);
}

// The epilogue's view of a layer pass, held in REGISTERS.  The persistent trunk kernel keeps the
// ConvParams of the current pass in shared memory; reading its fields through that copy cost one
// dependent LDS (+ branch) per field per row, re-issued after every mbarrier wait (asm memory
// clobber) -- most of the 1500 cycles a row's "math + stores" took (profiles/r1_fold_role_timing.txt).
// kTrunk: a residual-dense-block pass never upsamples and never writes the output frame; those
// fields become compile-time constants and their code disappears.
struct EpiRegs {
  const float* bias;
  int32_t cout, lrelu, c_off, level;
  const float* res1;
  const float* res2;
  float s1, s2;
  float* dst32a;
  float* dst32b;
  void* dst16;
  int32_t dst16_plane_px, dst16_coff, dst16_fmt, dst16_up;
  uint8_t* out_u8;
  int64_t out_stride, out_frame_stride;
  float* out_f32;
  int32_t out_h, out_w;
  int32_t out_trunc;
  int32_t debug_flags;
};
template <bool kTrunk>
__device__ __forceinline__ EpiRegs make_epi_regs(const ConvParams& p) {
  EpiRegs e;
  e.bias = p.bias; e.cout = p.cout; e.lrelu = p.lrelu; e.c_off = p.c_off; e.level = kTrunk ? 0 : p.level;
  e.res1 = p.res1; e.res2 = p.res2; e.s1 = p.s1; e.s2 = p.s2;
  e.dst32a = p.dst32a; e.dst32b = p.dst32b;
  e.dst16 = p.dst16; e.dst16_plane_px = p.dst16_plane_px; e.dst16_coff = p.dst16_coff; e.dst16_fmt = p.dst16_fmt;
  e.dst16_up = kTrunk ? 0 : p.dst16_up;
  e.out_u8 = kTrunk ? nullptr : p.out_u8; e.out_stride = p.out_stride; e.out_frame_stride = p.out_frame_stride;
  e.out_f32 = kTrunk ? nullptr : p.out_f32; e.out_h = p.out_h; e.out_w = p.out_w; e.out_trunc = p.out_trunc;
  e.debug_flags = dbg_flags(p);
  return e;
}

// Channels [n0, n0+16) of one pixel.  r1pre / r2pre: residual values the caller already fetched
// (the fold kernel issues those loads before it waits for the accumulator), or null.
// Params is ConvParams (kernel parameter space) or EpiRegs (registers).
template <class Params>
__device__ __forceinline__ void epilogue16(const Params& p, const TileGeom& tg, const PixelRef& px, int n0,
                                           float (&v)[16], const float* r1pre = nullptr, const float* r2pre = nullptr) {
  if (p.bias) {                                              // null: the caller has added the bias already
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = b4[i];                                  // global or shared memory (conv3x3_fold.cu stages the layer's biases)
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (p.lrelu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.f ? v[i] : 0.2f * v[i];
  }
  const int ch0 = p.c_off + n0;
  if (p.res1) {
    float r[16];
    if (!r1pre) { load_trunk16(p.res1, px.P, ch0, r); r1pre = r; }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], p.s1, r1pre[i]);
  }
  if (p.res2) {
    float r[16];
    if (!r2pre) { load_trunk16(p.res2, px.P, ch0, r); r2pre = r; }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], p.s2, r2pre[i]);
  }
  if (p.dst32a && !((NESR_PROF ? p.debug_flags : 0) & 4096)) {                 // 4096: no fp32 trunk stores (timing experiments)
    stg256f_stream(p.dst32a + trunk_offset(px.P, ch0), v);
    stg256f_stream(p.dst32a + trunk_offset(px.P, ch0 + 8), v + 8);
  }
  if (p.dst32b) {
    stg256f_stream(p.dst32b + trunk_offset(px.P, ch0), v);
    stg256f_stream(p.dst32b + trunk_offset(px.P, ch0 + 8), v + 8);
  }
  if (p.dst16 && !((NESR_PROF ? p.debug_flags : 0) & 16)) {                    // 16: no 16-bit activation stores (timing experiments)
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack2(v[2 * i], v[2 * i + 1], p.dst16_fmt);
    const int dch = p.dst16_coff + n0;                                             // 16 channels, 32-byte aligned
    uint16_t* base = reinterpret_cast<uint16_t*>(p.dst16) + static_cast<size_t>(dch >> 6) * p.dst16_plane_px * 64 + (dch & 63);
    if (!p.dst16_up) {
      stg256(base + static_cast<size_t>(px.P) * 64, w);
    } else {
      const LevelGeom g2 = tg.lv[p.level + 1];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const size_t P2 = static_cast<size_t>(g2.base) + static_cast<size_t>(2 * px.y + a) * g2.pitch + (2 * px.x + b);
          stg256(base + P2 * 64, w);
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
            if (p.out_trunc) {                                              // reference HEAD: clip(out * 255, 0, 255).astype(uint8)
              o[2 - c] = static_cast<uint8_t>(fminf(fmaxf(__fmul_rn(v[c], 255.f), 0.f), 255.f));
            } else {
              const float cl = fminf(fmaxf(v[c], 0.f), 1.f);
              o[2 - c] = static_cast<uint8_t>(rintf(__fmul_rn(cl, 255.f)));   // RGB -> BGR, half-to-even
            }
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
