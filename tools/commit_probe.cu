// commit_probe.cu -- what does a tcgen05.commit cost?  (run on a B200 via gpurun)
//
// The row-folded conv kernel issues one commit per ring stage plus one per output row.  With every
// other piece of work disabled the trunk still took ~420 cycles per commit (gpurun_out decompose
// traces), which suggests commits do not pipeline.  This probe measures it in isolation:
//
//   R repetitions of [ K MMAs (M=128, N=96, K=16, SS) ; C commits to distinct mbarriers ]
//   then one final commit + wait.  Reported: cycles per repetition, for K in {0,4,12,24,36}, C in {0,1,2,4}
//   and for two spellings of the commit (shared::cluster address vs generic address).
//
// A second experiment has a second warp WAIT on the committed barriers (as the real pipeline does),
// to see whether the cost is in the commit itself or in the wake-up of the waiter.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/commit_probe tools/commit_probe.cu
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../neural_enhanced_super_resolution_b200/csrc/ptx.cuh"

using namespace nesr;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);       \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ void commit_generic(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"(reinterpret_cast<uint64_t>(bar)) : "memory");
}

constexpr int kNBar = 16;

// mode 0: commits via shared::cluster address; mode 1: generic address; mode 2: plain mbarrier.arrive instead of commit
template <int MODE>
__global__ void __launch_bounds__(128, 1) commit_cost_kernel(int n, int kmma, int ncommit, int reps, int waiter, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABytes = 32 * 1024, kBBytes = 256 * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + kBBytes);       // kNBar ring + 1 final
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + kNBar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i <= kNBar; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(1, n);
    const uint32_t hi = umma_desc_hi_sw128();
    const uint32_t a_lo = umma_desc_lo(smem_u32(sa)), b_lo = umma_desc_lo(smem_u32(sb));
    __syncwarp();
    const long long t0 = clock64();
    int bi = 0;
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
        for (int k = 0; k < kmma; k += 4) umma_f16_ksteps<4>(tmem, a_lo + (k & 7) * 8, b_lo, hi, idesc);
      }
      __syncwarp();
      for (int c = 0; c < ncommit; ++c) {
        if (elect_one()) {
          if (MODE == 0) umma_commit(&bars[bi]);
          else if (MODE == 1) commit_generic(&bars[bi]);
          else mbar_arrive(&bars[bi]);
        }
        __syncwarp();
        bi = (bi + 1) % kNBar;
      }
    }
    if (elect_one()) umma_commit(&bars[kNBar]);
    __syncwarp();
    mbar_wait(&bars[kNBar], 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  } else if (warp == 1 && waiter) {
    // a consumer that waits for every committed barrier in order, like the TMA producer / epilogue do
    int bi = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r)
      for (int c = 0; c < ncommit; ++c) {
        for (int tries = 0; tries < 4000 && !mbar_try_wait(&bars[bi], ph); ++tries) {}   // bounded: a lapped waiter moves on
        if (++bi == kNBar) { bi = 0; ph ^= 1; }
      }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE>
double run(int n, int kmma, int ncommit, int waiter, int sms, long long* dc, std::vector<long long>& hc) {
  const int smem = 32 * 1024 + 256 * 128 + (kNBar + 1) * 8 + 64 + 1024;
  CK(cudaFuncSetAttribute(commit_cost_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int reps = 64;
  for (int it = 0; it < 2; ++it) commit_cost_kernel<MODE><<<sms, 128, smem>>>(n, kmma, ncommit, reps, waiter, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  CK(cudaMemcpy(hc.data(), dc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = hc[i] > mx ? hc[i] : mx;
  return (double)mx / reps;
}

int main() {
  CK(cudaSetDevice(0));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* dc;
  CK(cudaMalloc(&dc, sms * sizeof(long long)));
  std::vector<long long> hc(sms);
  const char* names[3] = {"commit shared::cluster", "commit generic addr", "plain mbarrier.arrive"};
  for (int waiter = 0; waiter < 2; ++waiter)
    for (int mode = 0; mode < 3; ++mode) {
      printf("== %s, N=96, waiter warp %s: cycles per repetition of [K MMAs ; C commits] ==\n", names[mode], waiter ? "ON" : "off");
      printf("   K\\C        0        1        2        4\n");
      for (int k : {0, 4, 12, 24, 36}) {
        printf("  %3d ", k);
        for (int c : {0, 1, 2, 4}) {
          double v = mode == 0 ? run<0>(96, k, c, waiter, sms, dc, hc) : mode == 1 ? run<1>(96, k, c, waiter, sms, dc, hc) : run<2>(96, k, c, waiter, sms, dc, hc);
          printf(" %8.1f", v);
        }
        printf("\n");
      }
    }
  printf("commit probe done\n");
  return 0;
}
