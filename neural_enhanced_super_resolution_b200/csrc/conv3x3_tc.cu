// conv3x3_tc.cu -- 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores.
//
//   D[128 pixels x N] += sum over 9 taps, Cin/16 k-steps of  A_tap[128 px x 16 ch] * W_tap[N x 16 ch]^T
//
// * A (activations) and B (weights) are staged in shared memory by TMA (cp.async.bulk.tensor,
//   128-byte swizzle) and consumed by tcgen05.mma (kind::f16, M=128, N in {16,32,64}, K=16) issued by
//   ONE thread; the fp32 accumulator lives in TMEM and is double-buffered so the epilogue of one
//   M-block overlaps the MMAs of the next.
// * thanks to the flat zero-padded layout (layout.h) the A tile of tap (dy,dx) is the 2-D box at
//   pixel coordinate  px + dy*pitch + dx  of the SAME tensor map: zero padding, tile borders and
//   row wrap need no special cases, and the dense-block concat is the channel coordinate.
// * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue
//   (tcgen05.ld -> bias/LeakyReLU/residual -> channel-offset / upsample-replicated / u8 stores).
// * persistent: grid = #SMs, static round-robin over the M-blocks of every tile in the batch.
#include <stdio.h>

#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace nesr {

namespace {

constexpr int kThreads = 192;
constexpr int kABytes = kBlockPixels * 128;                 // 128 px x 64 ch x 2 B

template <int N>
struct TcConfig {
  static constexpr int kBBytes = N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;     // multiple of 1024 for N in {16,32,64}
  static constexpr int kStages = (N == 64) ? 8 : (N == 32 ? 9 : 10);
  static constexpr int kTmemCols = (2 * N < 32) ? 32 : 2 * N;
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;   // + alignment slack
};

template <int N>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap,
                  const ConvParams p) {
  using Cfg = TcConfig<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&wmap);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nchunk = (p.cin + kChunkChannels - 1) / kChunkChannels;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = blockIdx.x; i < p.nblk; i += gridDim.x) {
        const BlockRef b = p.blocks[i];
        const int pitch = p.tiles[b.tile].lv[p.level].pitch;
        for (int c = 0; c < nchunk; ++c) {
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
            tma_load_2d(sa, &amap, &full[stage], 0, c * p.src_plane_px + b.px + (t / 3 - 1) * pitch + (t % 3 - 1));
            tma_load_2d(sa + kABytes, &wmap, &full[stage], 0, p.w_row0 + (t * nchunk + c) * N);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int i = blockIdx.x; i < p.nblk; i += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + as * N;
        uint32_t accumulate = 0;
        for (int c = 0; c < nchunk; ++c) {
          const int rem = (p.cin - c * kChunkChannels) >> 4;
          const int ksteps = rem < 4 ? rem : 4;
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
            const uint64_t ad = umma_smem_desc_sw128(a_addr, 1024);
            const uint64_t bd = umma_smem_desc_sw128(a_addr + kABytes, 1024);
            for (int k = 0; k < ksteps; ++k) {
              umma_f16(d, ad + 2 * k, bd + 2 * k, p.idesc, accumulate);      // +32 B along K per step
              accumulate = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ----------------------------------
    const int quarter = warp & 3;                 // TMEM lanes this warp may read
    const int row = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int i = blockIdx.x; i < p.nblk; i += gridDim.x) {
      const BlockRef b = p.blocks[i];
      const TileGeom& tg = p.tiles[b.tile];
      const PixelRef px = locate(tg.lv[p.level], b.px + row);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * N;
      uint32_t r[N / 16][16];
#pragma unroll
      for (int j = 0; j < N / 16; ++j) tmem_ld16(taddr + j * 16, r[j]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty[as]);                   // accumulator stage is free for the next M-block
      if (px.valid) {
#pragma unroll
        for (int j = 0; j < N / 16; ++j) {
          if (j * 16 < p.cout) {
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[j][e]);
            epilogue16(p, tg, px, j * 16, v);
          }
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int N>
cudaError_t launch_n(const CUtensorMap& amap, const CUtensorMap& wmap, const ConvParams& p, int num_sms,
                     cudaStream_t stream) {
  using Cfg = TcConfig<N>;
  const int grid = p.nblk < num_sms ? p.nblk : num_sms;
  conv3x3_tc_kernel<N><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(amap, wmap, p);
  return cudaGetLastError();
}

template <int N>
cudaError_t configure_n() {
  return cudaFuncSetAttribute(conv3x3_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TcConfig<N>::kSmemBytes);
}

}  // namespace

// Opt in to the large dynamic shared memory carve-out; once per device (engine create).
cudaError_t conv3x3_tc_configure() {
  cudaError_t e = configure_n<16>();
  if (e == cudaSuccess) e = configure_n<32>();
  if (e == cudaSuccess) e = configure_n<64>();
  return e;
}

cudaError_t launch_conv3x3_tc(const CUtensorMap& amap, const CUtensorMap& wmap, const ConvParams& p, int num_sms,
                              cudaStream_t stream) {
  if (p.nblk <= 0) return cudaSuccess;
  switch (p.npad) {
    case 16: return launch_n<16>(amap, wmap, p, num_sms, stream);
    case 32: return launch_n<32>(amap, wmap, p, num_sms, stream);
    case 64: return launch_n<64>(amap, wmap, p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nesr
