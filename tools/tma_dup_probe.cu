// Does cuTensorMapEncodeTiled accept a ZERO global stride, i.e. can TMA replicate pixels (nearest x2 along x) while it loads?
// 3-D view of a [pixels][64 ch] bf16 plane: (ch 64, dup 2 @ stride 0, px P @ 128 B); box (64, 2, 68) -> 136 slab rows = every source pixel twice.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_dup_probe tools/tma_dup_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__global__ void k(const __grid_constant__ CUtensorMap map, uint16_t* out, int px0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(136 * 128));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(d), "l"(reinterpret_cast<uint64_t>(&map)), "r"(b), "r"(0), "r"(0), "r"(px0) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < 136 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main() {
  cuInit(0);
  const int P = 1024;
  std::vector<uint16_t> h(P * 64);
  for (int p = 0; p < P; ++p) for (int c = 0; c < 64; ++c) h[p * 64 + c] = (uint16_t)(p * 64 + c);
  uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 136 * 64 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  for (int sw = 0; sw < 2; ++sw) {
    CUtensorMap m;
    const cuuint64_t gdim[3] = {64, 2, (cuuint64_t)P};
    const cuuint64_t gstride[2] = {0, 128};
    const cuuint32_t box[3] = {64, 2, 68}, estr[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %d: encode with zero stride -> CUresult %d\n", sw, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 128 + 1024);
    k<<<1, 128, 136 * 128>>>(m, o, 10);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> got(136 * 64);
    cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 136; ++row) for (int c = 0; c < 64; ++c) {
      int cc = c;
      if (sw) cc = (((c >> 3) ^ (row & 7)) << 3) | (c & 7);       // 128-byte swizzle: 16-byte chunk index XOR (row mod 8)
      const uint16_t want = (uint16_t)((10 + row / 2) * 64 + c);
      if (got[row * 64 + cc] != want) { if (bad < 4) printf("  row %d ch %d: got %u want %u\n", row, c, got[row * 64 + cc], want); ++bad; }
    }
    printf("  slab row r holds source pixel px0 + r/2: %s (%d mismatches)\n", bad ? "NO" : "YES", bad);
  }
  return 0;
}
