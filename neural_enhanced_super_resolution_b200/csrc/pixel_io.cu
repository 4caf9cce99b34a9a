// pixel_io.cu -- image -> level-0 network input, one pass:
//   BGR u8 HWC (or RGB fp32 NCHW) -> /255 -> RGB -> reflect pre-pad / mod-pad -> tile + halo gather
//   -> pixel_unshuffle(2) -> 16-bit channels 0..11 of the flat zero-padded [pixels][64] input buffer.
// Follows RealESRGANer.enhance / pre_process and basicsr pixel_unshuffle (oracle/realesrganer.py,
// oracle/rrdbnet.py):  feat[c*4 + i*2 + j][y][x] = img[c][2y+i][2x+j].
#include "epilogue.cuh"
#include "kernels.h"

namespace nesr {

namespace {

// torch 'reflect' padding on the right/bottom, applied twice (pre_pad, then mod pad).
__device__ __forceinline__ int unreflect(int s, int len, int pre_pad) {
  const int len1 = len + pre_pad;
  if (s >= len1) s = 2 * (len1 - 1) - s;
  if (s >= len) s = 2 * (len - 1) - s;
  return s;
}

__global__ void __launch_bounds__(kBlockPixels) pack_kernel(const PackParams p) {
  const BlockRef b = p.blocks[blockIdx.x];
  const TileGeom& tg = p.tiles[b.tile];
  const PixelRef px = locate(tg.lv[0], b.px + threadIdx.x);
  if (!px.valid) return;
  uint32_t packed[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int sy = unreflect(tg.src_y0 + 2 * px.y + i, p.H, p.pre_pad);
        const int sx = unreflect(tg.src_x0 + 2 * px.x + j, p.W, p.pre_pad);
        float f;
        if (p.in_f32_12) {                                   // channel c*4 + i*2 + j of the feature grid (H/2 x W/2), no padding
          f = p.in_f32_12[((static_cast<size_t>(tg.frame) * 12 + c * 4 + i * 2 + j) * (p.H >> 1) + (tg.src_y0 >> 1) + px.y) * (p.W >> 1) +
                          (tg.src_x0 >> 1) + px.x];
        } else if (p.in_u8) {
          const uint8_t u = p.in_u8[tg.frame * p.in_frame_stride + sy * p.in_stride + static_cast<int64_t>(sx) * 3 + (2 - c)];
          f = __fdiv_rn(static_cast<float>(u), 255.f);
        } else {
          f = p.in_f32[(static_cast<size_t>(tg.frame) * 3 + c) * p.H * p.W + static_cast<size_t>(sy) * p.W + sx];
        }
        v[i * 2 + j] = f;
      }
    packed[2 * c] = pack2(v[0], v[1], p.fmt);
    packed[2 * c + 1] = pack2(v[2], v[3], p.fmt);
  }
  uint2* d = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.x0) + static_cast<size_t>(px.P) * kChunkChannels);
  d[0] = make_uint2(packed[0], packed[1]);
  d[1] = make_uint2(packed[2], packed[3]);
  d[2] = make_uint2(packed[4], packed[5]);
}

}  // namespace

cudaError_t launch_pack(const PackParams& p, cudaStream_t stream) {
  if (p.nblk <= 0) return cudaSuccess;
  pack_kernel<<<p.nblk, kBlockPixels, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace nesr
