// conv3x3_body.cu -- the 23 x 3 residual dense blocks as ONE persistent kernel.
//
// Same row-folded tcgen05 / TMEM / TMA roles as conv3x3_fold.cu (fold_roles.cuh), but a CTA keeps its
// bands, its TMEM ring and its barriers across all 414 layer passes of the trunk (69 RDBs x [conv1..
// conv4, conv5 low half, conv5 high half], all with 32 output channels) and only swaps the resident
// folded weights between passes.  A pass may start once every CTA has finished the passes it depends
// on (ConvParams::sync_passes): the halo rows and columns of a band are written by neighbouring CTAs.
// That dependency is a grid-wide arrival counter in global memory -- release (fence + atomic) by one
// thread per CTA at the end of a pass, acquire-spin by the TMA producer before its first activation
// load of the next one; the weight load, which depends on nobody, is issued before the spin.  The two
// halves of conv5 read the same input and write disjoint channels, so no wait separates them.
//
// Versus one launch per layer pass this removes ~410 launches per frame, each of which cost a launch
// gap, a prologue (TMEM alloc, barrier init, descriptor fetch) and a cold pipeline -- together about a
// third of a small layer's wall time (profiles/r1_fold_role_timing.txt).
//
// Cross-proxy ordering: activations are written with generic-proxy stores and read back by TMA (async
// proxy), possibly by another CTA.  Writers execute fence.proxy.async before the CTA's release; the
// reader executes it after its acquire, before the first TMA load.
//
// Launched cooperatively (all CTAs co-resident: grid <= #SMs, 1 CTA/SM) so the spin cannot deadlock;
// every spin is bounded and traps instead of hanging.
#include "fold_roles.cuh"

namespace nesr {

namespace {

using namespace fold;

constexpr int COUT = 32;
using Cfg = FoldCfg<COUT>;
constexpr int kMaxChunks = kDense / kChunkChannels;                            // 3
constexpr int kBodyStages = 6;
constexpr int kWeightBytes = 3 * kMaxChunks * Cfg::kWBoxBytes;                 // 110592: largest pass (Cin 192)
constexpr int kSmemBytes = kWeightBytes + kBodyStages * kSlabBytes + kBarrierBytes + 1024;
static_assert(kSmemBytes <= kSmemBudget, "persistent trunk kernel: shared memory budget");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* ptr, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_body_kernel(const __grid_constant__ CUtensorMap amap_d0, const __grid_constant__ CUtensorMap amap_d1,
                    const __grid_constant__ CUtensorMap amap8_d0, const __grid_constant__ CUtensorMap amap8_d1,
                    const __grid_constant__ CUtensorMap wmap, const ConvParams* __restrict__ passes, const int npass,
                    unsigned* __restrict__ gbar) {
  __shared__ ConvParams sp;                                    // parameters of the current pass
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Pipe s = carve_pipe(smem, kWeightBytes, kBodyStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap_d0); tma_prefetch_desc(&amap_d1);
    tma_prefetch_desc(&amap8_d0); tma_prefetch_desc(&amap8_d1);
    tma_prefetch_desc(&wmap);
  }
  pipe_setup<COUT>(s, warp, lane, true);

  // pipeline positions live across passes (each thread belongs to exactly one role)
  RingPos rp;
  uint32_t u = 0;                                              // running row-slot counter of the TMEM ring

  for (int pass = 0; pass < npass; ++pass) {
    [[maybe_unused]] const long long prof_t0 = PROF_NOW();
    [[maybe_unused]] long long prof_t1 = 0, prof_t2 = 0, prof_t3 = 0;
    if (warp == 0) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(passes + pass);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&sp);
      for (int i = lane; i < static_cast<int>(sizeof(ConvParams) / 4); i += 32) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const ConvParams& p = sp;
    const int band_begin = p.cta_band_off[blockIdx.x];
    const int band_end = p.cta_band_off[blockIdx.x + 1];

    if (warp == 0) {
      load_weights<COUT>(p, &wmap, s);                         // depends on no other CTA
      if (p.sync_passes > 0) {                                 // every CTA has finished the first sync_passes passes
        if (lane == 0) {
          const unsigned target = static_cast<unsigned>(p.sync_passes) * gridDim.x;
          // relaxed spin, one acquire at the end: ld.acquire.gpu is LDG.STRONG + CCTL.IVALL, and an L1 invalidation per
          // poll makes every L1-cached load of the epilogue warps miss
          if (ld_relaxed_gpu(gbar) < target) {
            const long long t0 = clock64();
            while (ld_relaxed_gpu(gbar) < target) {
              if (clock64() - t0 > NESR_HANG_GUARD_CYCLES) __trap();
            }
          }
          (void)ld_acquire_gpu(gbar);
        }
        __syncwarp();
        fence_proxy_async_all();
      }
      prof_t1 = PROF_NOW();
      producer_bands(p, p.src_sel ? &amap_d1 : &amap_d0, p.src_sel ? &amap8_d1 : &amap8_d0, s, rp, band_begin, band_end);
      prof_t2 = PROF_NOW();
    } else if (warp == 1) {
      mma_bands<COUT>(p, s, rp, u, static_cast<uint32_t>(pass & 1), band_begin, band_end);
    } else {
      epilogue_bands<COUT, true>(p, s, u, warp, lane, band_begin, band_end);
      fence_proxy_async_all();                                 // generic-proxy stores -> later TMA (async proxy) reads
    }

    // end of pass: all MMAs have completed (the epilogue drained the last rows), the weights and `sp`
    // may be overwritten, and everything this CTA wrote is ordered before its arrival on the counter
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
      __threadfence();
      red_release_gpu_add(gbar, 1u);
    }
#if NESR_PROF
    if ((dbg_flags(sp) & 1024) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && pass < 48) {
      prof_t3 = PROF_NOW();
      printf("[body blk %d pass %d cin=%d] params+wload+spin %lld  producer %lld  drain+sync %lld  total %lld\n", (int)blockIdx.x, pass,
             sp.cin, prof_t1 - prof_t0, prof_t2 - prof_t1, prof_t3 - prof_t2, prof_t3 - prof_t0);
    }
    __syncthreads();                 // sp.debug_flags above must be read before the next pass overwrites sp
#endif
  }

  pipe_teardown(s, warp);
}

}  // namespace

cudaError_t conv3x3_body_configure() {
  return cudaFuncSetAttribute(conv3x3_body_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t launch_conv3x3_body(const CUtensorMap& d0, const CUtensorMap& d1, const CUtensorMap& d0_8, const CUtensorMap& d1_8,
                                const CUtensorMap& wmap, const ConvParams* d_passes, int npass, unsigned* d_gbar, int grid,
                                cudaStream_t stream) {
  if (grid <= 0 || npass <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_gbar, 0, sizeof(unsigned), stream);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;                 // co-residency guarantee for the arrival counter
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv3x3_body_kernel, d0, d1, d0_8, d1_8, wmap, d_passes, npass, d_gbar);
}

}  // namespace nesr
