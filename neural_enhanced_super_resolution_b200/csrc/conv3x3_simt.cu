// conv3x3_simt.cu -- CUDA-core 3x3 conv over exactly the buffers, packed weights and epilogue of the
// tcgen05 kernel.  TEST-ONLY cross-check (conv_impl = 1 / nesr_b200_debug_conv): it lets the GPU
// tests separate "tensor-core path wrong" from "layout / epilogue / host plan wrong".  It is never
// selected implicitly and is not a fallback: the product path is conv3x3_tc.cu.
#include "epilogue.cuh"
#include "kernels.h"

namespace nesr {

namespace {

__global__ void __launch_bounds__(kBlockPixels) conv3x3_simt_kernel(const ConvParams p) {
  const BlockRef b = p.blocks[blockIdx.x];
  const TileGeom& tg = p.tiles[b.tile];
  const LevelGeom g = tg.lv[p.level];
  const PixelRef px = locate(g, b.px + threadIdx.x);
  if (!px.valid) return;
  const int n0 = blockIdx.y * 16;
  if (n0 >= p.cout) return;
  const int nchunk = (p.cin + kChunkChannels - 1) / kChunkChannels;
  const int fp16 = p.fmt;                       // NESR_FMT_FP16 == 1
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int t = 0; t < 9; ++t) {
    const long long ps = static_cast<long long>(px.P) + (t / 3 - 1) * g.pitch + (t % 3 - 1);
    for (int ch = 0; ch < p.cin; ++ch) {
      const int chunk = ch >> 6, k = ch & 63;
      const float a = load16(p.src, (static_cast<size_t>(chunk) * p.src_plane_px + static_cast<size_t>(ps)) * 64 + k, fp16);
      const size_t wrow = static_cast<size_t>(p.w_row0) + static_cast<size_t>(t * nchunk + chunk) * p.npad + n0;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(a, load16(p.wpack, (wrow + j) * 64 + k, fp16), acc[j]);
    }
  }
  epilogue16(p, tg, px, n0, acc);
}

}  // namespace

cudaError_t launch_conv3x3_simt(const ConvParams& p, cudaStream_t stream) {
  if (p.nblk <= 0) return cudaSuccess;
  dim3 grid(p.nblk, p.npad / 16);
  conv3x3_simt_kernel<<<grid, kBlockPixels, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace nesr
