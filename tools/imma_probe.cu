// Issue rate of the legacy tensor path on sm_100a: mma.sync m16n8k32 u8 (IMMA.16832.U8.U8) against m16n8k16 f16 (HMMA.16816.F32), dependent
// accumulator chains of 4 independent MMAs per warp, 8 warps per CTA, one CTA per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -o imma_probe tools/imma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_imma(int* out, int iters) {
  int d[4][4] = {};
  uint32_t a[4] = {threadIdx.x, 2u, 3u, 4u}, b0 = threadIdx.x * 7u, b1 = 5u;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(d[j][0]), "+r"(d[j][1]), "+r"(d[j][2]), "+r"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  int s = 0;
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += d[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_hmma(float* out, int iters) {
  float d[4][4] = {};
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  float s = 0;
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += d[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int* o; cudaMalloc(&o, 148 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_imma<<<148, 256>>>(o, iters); else k_hmma<<<148, 256>>>((float*)o, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double n = 8.0 * 4 * iters;                      // MMAs per SM
      int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
      printf("%s: %.3f ms, %.1f ns per MMA per SM (%.1f cycles at %d MHz nominal), %.1f Tmac/s\n", which ? "HMMA.16816.F32 " : "IMMA.16832.U8 ", ms, ms * 1e6 / n,
             ms * 1e-3 * clk * 1e3 / n, clk / 1000, (which ? 2048.0 : 4096.0) * n * 148 / (ms * 1e-3) / 1e12);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
