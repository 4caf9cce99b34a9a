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

// The reference HEAD's 12-channel network input (nesr/nesr.py:859-880 `_apply_esrgan_12channel`, :916-925 `_3channel`), computed from the
// u8 image on the fly: t = BGR / 255 (the reference feeds BGR unswapped); channels 0-2 t, 3-5 clamp(t * 1.1, 0, 1), 6-8 clamp(t * 0.9, 0, 1),
// 9-11 cv2.GaussianBlur(img, (3, 3), 0) / 255 -- the fixed 1-2-1 kernel, BORDER_REFLECT_101, (sum + 8) >> 4 -- or four copies of t.  Same
// fp32 operations as the torch expressions (true division, multiplication by float32(1.1) / float32(0.9)).
__device__ __forceinline__ float head_channel(const PackParams& p, int k, int y, int x) {
  const int h = p.H >> 1, w = p.W >> 1;
  const int variant = p.head_replicate ? 0 : k / 3, color = 2 - k % 3;          // t[0] = B = rgb[2]
  auto at = [&](int yy, int xx) {
    yy = yy < 0 ? -yy : (yy >= h ? 2 * (h - 1) - yy : yy);
    xx = xx < 0 ? -xx : (xx >= w ? 2 * (w - 1) - xx : xx);
    yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);                                     // 1-pixel-wide images
    xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
    return static_cast<int>(p.in_u8_head[static_cast<int64_t>(yy) * p.in_stride + static_cast<int64_t>(xx) * 3 + color]);
  };
  if (variant == 3) {
    const int s = at(y - 1, x - 1) + 2 * at(y - 1, x) + at(y - 1, x + 1) + 2 * (at(y, x - 1) + 2 * at(y, x) + at(y, x + 1)) +
                  at(y + 1, x - 1) + 2 * at(y + 1, x) + at(y + 1, x + 1);
    return __fdiv_rn(static_cast<float>((s + 8) >> 4), 255.f);
  }
  const float t = __fdiv_rn(static_cast<float>(at(y, x)), 255.f);
  if (variant == 1) return fminf(fmaxf(__fmul_rn(t, 1.1f), 0.f), 1.f);
  if (variant == 2) return fminf(fmaxf(__fmul_rn(t, 0.9f), 0.f), 1.f);
  return t;
}

__global__ void __launch_bounds__(kBlockPixels) pack_kernel(const PackParams p) {
  const BlockRef b = p.blocks[blockIdx.x];
  const TileGeom& tg = p.tiles[b.tile];
  const PixelRef px = locate(tg.lv[0], b.px + threadIdx.x);
  if (!px.valid) return;
  if (p.in_f32_12 && p.feat_ch != 12) {                      // any channel count on the feature grid (scale-4 / scale-1 architectures)
    const size_t plane = static_cast<size_t>(p.H >> 1) * (p.W >> 1);
    const float* src = p.in_f32_12 + static_cast<size_t>(tg.frame) * p.feat_ch * plane +
                       static_cast<size_t>((tg.src_y0 >> 1) + px.y) * (p.W >> 1) + (tg.src_x0 >> 1) + px.x;
    uint32_t* d32 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.x0) + static_cast<size_t>(px.P) * kChunkChannels);
    for (int k = 0; k < p.feat_ch; k += 2)
      d32[k >> 1] = pack2(src[k * plane], k + 1 < p.feat_ch ? src[(k + 1) * plane] : 0.f, p.fmt);
    return;
  }
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
        if (p.in_u8_head) {                                  // channel c*4 + i*2 + j of the HEAD's 12-channel tensor, built here
          f = head_channel(p, c * 4 + i * 2 + j, (tg.src_y0 >> 1) + px.y, (tg.src_x0 >> 1) + px.x);
        } else if (p.in_f32_12) {                            // channel c*4 + i*2 + j of the feature grid (H/2 x W/2), no padding
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

namespace {
// One block per (tile, 8 output rows): copies the tile's rectangle from its slot to its place in the frame, 16 bytes per
// thread where both sides allow it (tile origins are multiples of tile_out * 3 bytes; edge tiles and odd strides fall back to bytes).
struct TileIds { int32_t id[kUnpackMaxTiles]; };     // slot k holds tile id[k] (< 0: empty slot); id[0] < -1: slots are the range first, first+1, ...
__global__ void __launch_bounds__(256) unpack_tiles_kernel(const uint8_t* __restrict__ slots, int slot_w, int slot_h, int tiles_x, int tile_out,
                                                           int out_h, int out_w, int first, const __grid_constant__ TileIds ids,
                                                           uint8_t* __restrict__ out, int64_t out_stride) {
  const int k = blockIdx.y, t = ids.id[0] < -1 ? first + k : ids.id[k];
  if (t < 0) return;
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int y0 = ty * tile_out, x0 = tx * tile_out;
  const int h = min(tile_out, out_h - y0), w = min(tile_out, out_w - x0);
  if (h <= 0 || w <= 0) return;
  const uint8_t* src = slots + static_cast<size_t>(k) * slot_w * 3 * slot_h;
  const int64_t row_bytes = static_cast<int64_t>(w) * 3;
  for (int r = blockIdx.x * 8; r < min(h, blockIdx.x * 8 + 8); ++r) {
    const uint8_t* s = src + static_cast<size_t>(r) * slot_w * 3;
    uint8_t* d = out + static_cast<int64_t>(y0 + r) * out_stride + static_cast<int64_t>(x0) * 3;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | static_cast<uintptr_t>(row_bytes)) & 15) == 0) {
      for (int64_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(d)[i] = __ldg(reinterpret_cast<const uint4*>(s) + i);
    } else {
      for (int64_t i = threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = s[i];
    }
  }
}
}  // namespace

cudaError_t launch_unpack_tiles(const uint8_t* slots, int slot_w, int slot_h, int tiles_x, int tile_out, int out_h, int out_w, int first,
                                int count, const int32_t* tile_ids, uint8_t* out, int64_t out_stride, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  const int rows = tile_out < out_h ? tile_out : out_h;
  const size_t slot_bytes = static_cast<size_t>(slot_w) * 3 * slot_h;
  for (int k0 = 0; k0 < count; k0 += kUnpackMaxTiles) {
    const int n = count - k0 < kUnpackMaxTiles ? count - k0 : kUnpackMaxTiles;
    TileIds ids;
    ids.id[0] = -2;
    if (tile_ids) for (int k = 0; k < n; ++k) ids.id[k] = tile_ids[k0 + k] < 0 ? -1 : tile_ids[k0 + k];
    unpack_tiles_kernel<<<dim3((rows + 7) / 8, n), 256, 0, stream>>>(slots + k0 * slot_bytes, slot_w, slot_h, tiles_x, tile_out, out_h, out_w,
                                                                      first + k0, ids, out, out_stride);
  }
  return cudaGetLastError();
}

}  // namespace nesr
