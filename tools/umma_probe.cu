// umma_probe.cu -- hardware probes that decide the conv kernel's design (run on a B200 via gpurun):
//
//  (1) SHIFTED-VIEW probe: can one shared-memory slab of pixels written by TMA (128-byte swizzle)
//      serve all 3x3 taps as tcgen05.mma A operands whose start address is shifted by a multiple
//      of 128 B (one pixel row), i.e. NOT 1024-byte aligned?  Tests start = slab + s*128 with the
//      descriptor's base_offset field = 0 and = (s & 7).
//  (2) THROUGHPUT probe: cycles per tcgen05.mma (M=128, K=16, operands in shared memory) for
//      N = 16..256 -- the SMEM-operand-bandwidth ceiling that bounds the N=32 / N=64 conv layers.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_probe tools/umma_probe.cu
#include <cuda_bf16.h>
#include <math.h>
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

constexpr int kSlabRows = 384;   // pixels in the slab (48 KB)
constexpr int kN = 32;

// One CTA: TMA-load the slab [kSlabRows x 64] and B [kN x 64], run 4 k-steps of M=128 MMA with the A
// start shifted by `shift_rows` rows, write D (128 x kN fp32) to global.
__global__ void __launch_bounds__(128, 1)
shift_probe_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int shift_rows,
                   int base_offset, int sbo_bytes, float* d_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + kSlabRows * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kN * 128);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, kSlabRows * 128 + kN * 128);
    for (int r = 0; r < kSlabRows; r += 128) tma_load_2d(sa + r * 128, &amap, bar, 0, r);
    tma_load_2d(sb, &bmap, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(1, kN);
    const uint64_t ad = umma_smem_desc_sw128(smem_u32(sa) + shift_rows * 128, sbo_bytes, base_offset);
    const uint64_t bd = umma_smem_desc_sw128(smem_u32(sb), 1024);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, k > 0);
    umma_commit(mbar);
  }
  __syncwarp();
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t r[16];
  for (int j = 0; j < kN / 16; ++j) {
    __syncwarp();
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + j * 16, r);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) d_out[(warp * 32 + lane) * kN + j * 16 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// Throughput: `iters` back-to-back MMAs (M=128, K=16) with N = n on operands resident in smem.
// mode 0: same A every time; mode 1: A start cycles over 9 row-shifted views (as the conv taps do).
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, int mode, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // 64 KB of A rows
  uint8_t* sb = smem + 65536;               // 256 rows x 128 B
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sb + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(1, n);
    const uint64_t bd = umma_smem_desc_sw128(smem_u32(sb), 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int view = mode ? (i % 9) : 0;
      const uint64_t ad = umma_smem_desc_sw128(smem_u32(sa) + view * 2048, 1024) + 2 * (i & 3);
      umma_f16(tmem, ad, bd + 2 * (i & 3), idesc, 1);
    }
    umma_commit(mbar);
    mbar_wait(mbar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* base, int rows, int box_rows) {
  CUtensorMap m;
  const cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  return m;
}

int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;

  // ---- (1) shifted views ----
  std::vector<__nv_bfloat16> ha(kSlabRows * 64), hb(kN * 64);
  std::vector<float> fa(kSlabRows * 64), fb(kN * 64);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 9 - 4) / 4.0f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db;
  float* dd;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dd, 128 * kN * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap amap = make_map(enc, da, kSlabRows, 128), bmap = make_map(enc, db, kN, kN);
  const int smem1 = kSlabRows * 128 + kN * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
  // expected rows for every possible source row
  std::vector<float> exp_rows((kSlabRows) * kN);
  for (int r = 0; r < kSlabRows; ++r)
    for (int n = 0; n < kN; ++n) {
      float acc = 0;
      for (int k = 0; k < 64; ++k) acc += fa[r * 64 + k] * fb[n * 64 + k];
      exp_rows[r * kN + n] = acc;
    }
  std::vector<float> hd(128 * kN);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 17, 131, 200};
  printf("== shifted-view probe: A start = slab + s*128 B, SBO=1024 ==\n");
  for (int variant = 0; variant < 2; ++variant)
    for (int s : shifts) {
      const int bo = variant == 0 ? 0 : (s & 7);
      CK(cudaMemset(dd, 0, 128 * kN * 4));
      shift_probe_kernel<<<1, 128, smem1>>>(amap, bmap, s, bo, 1024, dd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("s=%d base_offset=%d : kernel failed: %s\n", s, bo, cudaGetErrorString(e)); return 3; }
      CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
      int good = 0;
      int map16[16];
      for (int m = 0; m < 128; ++m) {
        bool ok = true;
        for (int n = 0; n < kN; ++n) ok &= fabsf(hd[m * kN + n] - exp_rows[(s + m) * kN + n]) < 1e-3f;
        good += ok;
        if (m < 16) {   // which source row does output row m actually hold?
          map16[m] = -1;
          for (int r = 0; r < kSlabRows && map16[m] < 0; ++r) {
            bool eq = true;
            for (int n = 0; n < kN; ++n) eq &= fabsf(hd[m * kN + n] - exp_rows[r * kN + n]) < 1e-3f;
            if (eq) map16[m] = r;
          }
        }
      }
      printf("s=%3d base_offset=%d : %3d/128 rows correct%s | rows0-15 <- src", s, bo, good, good == 128 ? "  OK" : "  MISMATCH");
      for (int m = 0; m < 16; ++m) printf(" %d", map16[m]);
      printf("\n");
    }

  // ---- (2) MMA issue rate vs N ----
  printf("== tcgen05.mma rate, M=128 K=16, SS operands, 1 CTA/SM on all SMs ==\n");
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* dc;
  CK(cudaMalloc(&dc, sms * sizeof(long long)));
  const int smem2 = 65536 + 32768 + 64 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  std::vector<long long> hc(sms);
  const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {1, sms})
      for (int n : ns) {
        const int iters = 8192;
        mma_rate_kernel<<<grid, 128, smem2>>>(n, iters, mode, dc);   // warm
        mma_rate_kernel<<<grid, 128, smem2>>>(n, iters, mode, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(hc.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
        const double cyc = (double)mx / iters;
        const double macs = 128.0 * n * 16 / cyc;
        printf("mode=%d grid=%3d N=%3d : %.2f cyc/MMA  %.0f MAC/clk/SM  (%.1f%% of 4096)  smem operand B/clk=%.0f\n", mode, grid, n,
               cyc, macs, 100.0 * macs / 4096.0, (4096.0 + n * 32.0) / cyc);
      }
  printf("probe done\n");
  return 0;
}
