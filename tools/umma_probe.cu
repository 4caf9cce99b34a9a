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

// Throughput: MMAs (M=128, K=16, N=n) on operands resident in smem, issue loop fully unrolled so the
// issuing thread does ~2 integer instructions per MMA (descriptor = base + immediate).
//   NACC  : number of independent TMEM accumulators the MMAs rotate over
//   ORDER : 0 = accumulator index fastest (consecutive MMAs hit DIFFERENT accumulators)
//           1 = k-step, then tap, then accumulator (36 consecutive MMAs chain on the SAME accumulator)
//   AVIEW : 0 = one A view / one B view; 2 = nine A views shifted by the flat 3x3 tap offsets of a
//           pitch-267 tile (NOT 1024-B aligned) and nine B (weight) views -- the conv's real pattern
__device__ __forceinline__ constexpr int tap_rows(int t) { return (t / 3) * 267 + (t % 3); }

template <int NACC, int ORDER, int AVIEW>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int reps, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABytes = 96 * 1024, kBBytes = 9 * 32 * 128 + 256 * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + kABytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sb + kBBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(1, n);
    const uint64_t a0 = umma_smem_desc_sw128(smem_u32(sa), 1024);
    const uint64_t b0 = umma_smem_desc_sw128(smem_u32(sb), 1024);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (ORDER == 0) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int acc = 0; acc < NACC; ++acc) {
              const int rows = (AVIEW ? tap_rows(tap) : 0) + acc * 8;
              umma_f16(tmem + acc * n, a0 + (rows * 128 >> 4) + 2 * k, b0 + (AVIEW ? (tap * 32 * 128 >> 4) : 0) + 2 * k, idesc, 1);
            }
      } else {
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rows = (AVIEW ? tap_rows(tap) : 0) + acc * 8;
              umma_f16(tmem + acc * n, a0 + (rows * 128 >> 4) + 2 * k, b0 + (AVIEW ? (tap * 32 * 128 >> 4) : 0) + 2 * k, idesc, 1);
            }
      }
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
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NACC, int ORDER, int AVIEW>
void run_rate(int n, int sms, long long* dc, std::vector<long long>& hc) {
  if (NACC * n > 512) return;
  const int smem2 = 96 * 1024 + 9 * 32 * 128 + 256 * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel<NACC, ORDER, AVIEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  const int reps = 16;
  mma_rate_kernel<NACC, ORDER, AVIEW><<<sms, 128, smem2>>>(n, reps, dc);
  mma_rate_kernel<NACC, ORDER, AVIEW><<<sms, 128, smem2>>>(n, reps, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  CK(cudaMemcpy(hc.data(), dc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = hc[i] > mx ? hc[i] : mx;
  const double cyc = (double)mx / (reps * 36 * NACC);
  const double macs = 128.0 * n * 16 / cyc;
  printf("aview=%d order=%d N=%3d nacc=%2d : %6.2f cyc/MMA  %4.0f MAC/clk/SM (%5.1f%%)  smem operand B/clk=%.0f\n", AVIEW, ORDER, n, NACC, cyc,
         macs, 100.0 * macs / 4096.0, (4096.0 + n * 32.0) / cyc);
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

  // ---- (2) MMA rate vs N, accumulator interleave and A-view pattern ----
  printf("== tcgen05.mma rate, M=128 K=16, SS operands, 1 CTA/SM on all SMs ==\n");
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* dc;
  CK(cudaMalloc(&dc, sms * sizeof(long long)));
  std::vector<long long> hc(sms);
  for (int n : {16, 32, 64, 96, 128, 192, 256}) {
    run_rate<1, 0, 0>(n, sms, dc, hc);
    run_rate<2, 0, 0>(n, sms, dc, hc);
    run_rate<4, 0, 0>(n, sms, dc, hc);
    run_rate<8, 0, 0>(n, sms, dc, hc);
    run_rate<4, 1, 0>(n, sms, dc, hc);
    run_rate<8, 1, 0>(n, sms, dc, hc);
  }
  for (int n : {16, 32, 64}) {            // B views are 32 rows apart: N <= 64 keeps them inside the region
    run_rate<1, 0, 2>(n, sms, dc, hc);
    run_rate<2, 0, 2>(n, sms, dc, hc);
    run_rate<4, 0, 2>(n, sms, dc, hc);
    run_rate<8, 0, 2>(n, sms, dc, hc);
    run_rate<4, 1, 2>(n, sms, dc, hc);
    run_rate<8, 1, 2>(n, sms, dc, hc);
  }
  printf("probe done\n");
  return 0;
}
