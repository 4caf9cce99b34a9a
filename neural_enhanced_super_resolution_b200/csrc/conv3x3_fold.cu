// conv3x3_fold.cu -- one 3x3 conv layer pass per launch: row-folded implicit GEMM on tcgen05 / TMEM / TMA.
//
// The roles (TMA producer, MMA issuer, epilogue warps) and the reasoning behind the row fold live in
// fold_roles.cuh.  This kernel runs the six edge layers of the network (conv_first, conv_body,
// conv_up1/2, conv_hr, conv_last); the 69 residual dense blocks in between run inside the persistent
// kernels (conv3x3_trunk.cu; conv3x3_body.cu, which uses the same roles).  (conv_impl = 3 runs every layer through this
// kernel -- the previous production path, kept as a cross-check.)
//
// Layers with Cout = 64 and Cin > 64 (RDB conv5) run as two passes of 32 output channels so that
// the resident weights (3*nchunk boxes of [96 x 64]) fit next to the activation ring.
#include "fold_roles.cuh"

namespace nesr {

namespace {

using namespace fold;

template <int COUT>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_fold_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap amap8,
                    const __grid_constant__ CUtensorMap wmap, const ConvParams p) {
  using Cfg = FoldCfg<COUT>;
#if NESR_PROF
  unsigned long long prof_cta_start; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_cta_start));
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunk = (p.cin + kChunkChannels - 1) / kChunkChannels;
  Pipe s = carve_pipe(smem, 3 * nchunk * Cfg::kWBoxBytes, p.fold_stages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&amap8);
    tma_prefetch_desc(&wmap);
  }
  if (threadIdx.x < 64) s.bias[threadIdx.x] = threadIdx.x < COUT ? p.bias[threadIdx.x] : 0.f;   // constants of the network: no need to wait for the previous layer
  pipe_setup<COUT>(s, warp, lane, !(dbg_flags(p) & 128));

  const int band_begin = p.cta_band_off[blockIdx.x];
  const int band_end = p.cta_band_off[blockIdx.x + 1];
  pdl_launch_dependents();        // the next layer may start its prologue on SMs this grid has left

  if (warp == 0) {
    load_weights<COUT>(p, &wmap, s);
    pdl_wait();                    // activations below were written by the previous layer
    RingPos rp;
    producer_bands(p, &amap, &amap8, s, rp, band_begin, band_end);
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {             // single-thread issuer (fold_roles.cuh mma_bands, kSingle)
      RingPos rp;
      uint32_t u = 0;
      mma_bands<COUT, true>(p, s, rp, u, 0, band_begin, band_end);
    }
    __syncwarp();
  } else {
    pdl_wait();                    // residual reads / stores must not overtake the previous layer
    uint32_t u = 0;
    epilogue_bands<COUT, false>(p, s, u, warp, lane, band_begin, band_end);
  }

  pipe_teardown(s, warp);
#if NESR_PROF
  if ((dbg_flags(p) & 256) && threadIdx.x == 0) {       // per-CTA lifetime: who are the stragglers?
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    int nrows = 0;
    for (int bi = band_begin; bi < band_end; ++bi) nrows += p.bands[bi].rows + 2;
    printf("[cta cin=%d cout=%d] blk %d sm %u start_ns %llu end_ns %llu bands %d inrows %d\n", p.cin, COUT, (int)blockIdx.x, smid,
           prof_cta_start, t1, band_end - band_begin, nrows);
  }
#endif
}

template <int COUT>
int stages_for(int nchunk) {
  const int wbytes = 3 * nchunk * FoldCfg<COUT>::kWBoxBytes;
  int s = (kSmemBudget - 1024 - kBarrierBytes - wbytes) / kSlabBytes;
  return s > kMaxStages ? kMaxStages : s;
}

template <int COUT>
cudaError_t launch_c(const CUtensorMap& amap, const CUtensorMap& amap8, const CUtensorMap& wmap, ConvParams p, int grid,
                     cudaStream_t stream) {
  const int nchunk = (p.cin + kChunkChannels - 1) / kChunkChannels;
  p.fold_stages = stages_for<COUT>(nchunk);
  if (p.fold_stages < 2) return cudaErrorInvalidConfiguration;
  const int smem = 3 * nchunk * FoldCfg<COUT>::kWBoxBytes + p.fold_stages * kSlabBytes + kBarrierBytes + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (dbg_flags(p) & 512) ? 0 : 1;      // 512: no programmatic dependent launch (timing experiments)
  return cudaLaunchKernelEx(&cfg, conv3x3_fold_kernel<COUT>, amap, amap8, wmap, p);
}

template <int COUT>
cudaError_t configure_c() {
  return cudaFuncSetAttribute(conv3x3_fold_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

}  // namespace

cudaError_t conv3x3_fold_configure() {
  cudaError_t e = configure_c<16>();
  if (e == cudaSuccess) e = configure_c<32>();
  if (e == cudaSuccess) e = configure_c<64>();
  return e;
}

// Can a (cin16, npad) layer pass keep its folded weights resident with at least 3 ring stages?
bool conv3x3_fold_fits(int cin16, int npad) {
  const int nchunk = (cin16 + kChunkChannels - 1) / kChunkChannels;
  switch (npad) {
    case 16: return stages_for<16>(nchunk) >= 3;
    case 32: return stages_for<32>(nchunk) >= 3;
    case 64: return stages_for<64>(nchunk) >= 3;
    default: return false;
  }
}

cudaError_t launch_conv3x3_fold(const CUtensorMap& amap, const CUtensorMap& amap8, const CUtensorMap& wmap,
                                const ConvParams& p, int grid, cudaStream_t stream) {
  if (grid <= 0) return cudaSuccess;
  switch (p.npad) {
    case 16: return launch_c<16>(amap, amap8, wmap, p, grid, stream);
    case 32: return launch_c<32>(amap, amap8, wmap, p, grid, stream);
    case 64: return launch_c<64>(amap, amap8, wmap, p, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nesr
