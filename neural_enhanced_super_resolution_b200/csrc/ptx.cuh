// ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the conv kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the UMMA
// shared-memory + instruction descriptors.  sm_100a only; no fallbacks.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef NESR_PROF
#define NESR_PROF 0     // 1: build with device printf + per-role cycle accounting (NESR_B200_PROF=1 python -m ..._build)
#endif

namespace nesr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp.  The async-proxy instructions (TMA, tcgen05.mma/commit) take
// their operands from UNIFORM registers: code that runs warp-wide with uniform values and elects a
// lane only around the instruction itself compiles to plain UMOV/R2UR, whereas code nested under
// `if (lane == 0)` gets a per-instruction ELECT/R2UR/BRA "waterfall" loop (seen in SASS: ~17
// instructions per MMA), which made the issuing thread the bottleneck of the first fold kernel.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: a kernel launched with programmaticStreamSerialization may start
// (prologue, weight loads) while its predecessor in the stream is still draining; pdl_wait() blocks
// until the predecessor has completed and its writes are visible, pdl_launch_dependents() lets the
// successor begin as soon as this grid's CTAs free their SMs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking look: mbarrier.try_wait suspends the thread for a hardware time limit before it answers "not yet"; test_wait answers at once.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU box.
#ifndef NESR_HANG_GUARD_CYCLES
#define NESR_HANG_GUARD_CYCLES (4000000000LL)   // ~2-3 s at B200 clocks
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > NESR_HANG_GUARD_CYCLES) {
#if NESR_PROF
      printf("nesr_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
#endif
      __trap();                     // surfaces as a failed launch (cudaErrorLaunchFailure), never a hang
    }
  }
}

// ------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tiled load: coordinates are (innermost, outer); out-of-bounds elements are zero-filled.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 3-D tiled load.  Used with a ZERO-stride middle dimension (channels, 2 copies @ stride 0, pixels): the TMA engine writes every source
// pixel twice, i.e. it delivers the nearest-x2 up-sampled row (tools/tma_dup_probe.cu: the driver accepts the stride, also with the
// 128-byte swizzle) -- the read-side form of upstream's F.interpolate(scale_factor=2, mode='nearest') for conv_up1 / conv_up2.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 eviction policies.  The dense-block activations of a tile group are re-read by every layer pass and should
// stay in the 126 MB L2 (evict_last); the fp32 trunk streams through once per dense block and must not push them
// out (evict_first on those loads / stores, see epilogue.cuh).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// The same load delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA in `mask`.
__device__ __forceinline__ void tma_load_2d_hint_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                                    uint16_t mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask), "l"(policy)
      : "memory");
}

// Fire-and-forget prefetch of a 2-D box into L2 (no shared-memory destination, no barrier).
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1)
               : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile whose rows are 128 bytes (64 16-bit
// elements) written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: 8-row groups of 1024 B, group stride
// (SBO) `sbo_bytes`.  bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [49,52) base offset, [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                              // LBO: unused for swizzled K-major
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (sm_100)
  d |= static_cast<uint64_t>(base_offset & 7u) << 49;
  d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: fp32 accumulate, A and B both `fmt` (0 = fp16, 1 = bf16 in
// the hardware encoding), both K-major, M = 128, N = n.
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t hw_fmt, uint32_t n) {
  return (1u << 4) | (hw_fmt << 7) | (hw_fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// KS back-to-back accumulating MMAs (k-steps of one 64-channel chunk) from ONE asm block: the
// descriptors differ only in the low word (+32 B = +2 per k-step), so the issuing thread spends a
// handful of integer instructions per MMA instead of rebuilding 64-bit descriptors.
// lo = (addr >> 4) | LBO field; hi = SBO | version | swizzle (same for A and B here).
__device__ __forceinline__ constexpr uint32_t umma_desc_hi_sw128(uint32_t sbo_bytes = 1024) {
  return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
// K-major operand written by TMA with CU_TENSOR_MAP_SWIZZLE_64B: rows of 64 bytes (32 16-bit elements), 8-row groups of 512 B
__device__ __forceinline__ constexpr uint32_t umma_desc_hi_sw64(uint32_t sbo_bytes = 512) {
  return (sbo_bytes >> 4) | (1u << 14) | (4u << 29);
}
// Two k-steps of a half-width (64-byte-row) A operand against a 128-byte-row B operand: separate descriptor high words.
#define NESR_UMMA_STEP_AB(K)                                           \
  "add.u32 ta, %1, " #K ";\n\t"                                        \
  "add.u32 tb, %2, " #K ";\n\t"                                        \
  "mov.b64 da, {ta, %5};\n\t"                                          \
  "mov.b64 db, {tb, %3};\n\t"                                          \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
__device__ __forceinline__ void umma_f16_2ksteps_half_a(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi_b, uint32_t idesc, uint32_t hi_a) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               NESR_UMMA_STEP_AB(0) NESR_UMMA_STEP_AB(2) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi_b), "r"(idesc), "r"(hi_a) : "memory");
}

#define NESR_UMMA_STEP(K)                                              \
  "add.u32 ta, %1, " #K ";\n\t"                                        \
  "add.u32 tb, %2, " #K ";\n\t"                                        \
  "mov.b64 da, {ta, %3};\n\t"                                          \
  "mov.b64 db, {tb, %3};\n\t"                                          \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"

template <int KS>
__device__ __forceinline__ void umma_f16_ksteps(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  static_assert(KS >= 1 && KS <= 4, "one 64-channel chunk has at most 4 k-steps");
  if constexpr (KS == 1) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 NESR_UMMA_STEP(0) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  } else if constexpr (KS == 2) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 NESR_UMMA_STEP(0) NESR_UMMA_STEP(2) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  } else if constexpr (KS == 3) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 NESR_UMMA_STEP(0) NESR_UMMA_STEP(2) NESR_UMMA_STEP(4) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 NESR_UMMA_STEP(0) NESR_UMMA_STEP(2) NESR_UMMA_STEP(4) NESR_UMMA_STEP(6) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  }
}
__device__ __forceinline__ void umma_f16_ksteps_rt(int ks, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  switch (ks) {
    case 4: umma_f16_ksteps<4>(tmem_d, a_lo, b_lo, hi, idesc); break;
    case 2: umma_f16_ksteps<2>(tmem_d, a_lo, b_lo, hi, idesc); break;
    case 1: umma_f16_ksteps<1>(tmem_d, a_lo, b_lo, hi, idesc); break;
    default: umma_f16_ksteps<3>(tmem_d, a_lo, b_lo, hi, idesc); break;
  }
}

// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of one cluster drive both SMs' tensor cores with ONE tcgen05.mma (M = 256: each
// CTA supplies 128 rows of A and half of B's N rows from the same shared-memory offsets, and receives its 128 rows of D
// at the same TMEM address).  Only the leader (cluster rank 0) issues MMAs; barriers it waits on collect arrivals and TMA
// bytes from both CTAs through shared::cluster addresses; its commits are multicast to both CTAs' barriers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// RELAXED on purpose: a .release.cluster arrive compiles to MEMBAR.ALL.GPU + arrive (measured: +4 ms per tile group when
// every slab's arrive carried one).  The producers publish nothing through these arrives (the data arrives with the TMA's
// own complete_tx), and the epilogue's TMEM accesses are ordered by tcgen05.wait + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are signalled on a barrier of either CTA of the pair
__device__ __forceinline__ void tma_load_2d_hint_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int32_t c0,
                                                     int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {   // whole warp, in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of cta_group::2 MMAs, multicast to the barrier at the same offset in every CTA of `mask`
// cta_group::1 commit whose arrive lands on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
#define NESR_UMMA2_STEP(K)                                             \
  "add.u32 ta, %1, " #K ";\n\t"                                        \
  "add.u32 tb, %2, " #K ";\n\t"                                        \
  "mov.b64 da, {ta, %3};\n\t"                                          \
  "mov.b64 db, {tb, %3};\n\t"                                          \
  "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
template <int KS>
__device__ __forceinline__ void umma2_f16_ksteps(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  static_assert(KS == 2 || KS == 4, "trunk chunks have 2 or 4 k-steps");
  if constexpr (KS == 2) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 NESR_UMMA2_STEP(0) NESR_UMMA2_STEP(2) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 ta, tb;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 NESR_UMMA2_STEP(0) NESR_UMMA2_STEP(2) NESR_UMMA2_STEP(4) NESR_UMMA2_STEP(6) "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  }
}
// M = 256 instruction descriptor (cta_group::2), kind::f16, fp32 accumulate, K-major A and B
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16_m256(uint32_t hw_fmt, uint32_t n) {
  return (1u << 4) | (hw_fmt << 7) | (hw_fmt << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

// TMEM -> registers: each thread of the warp reads 16 consecutive 32-bit columns of its own lane
// (warp w of the CTA may only touch lanes 32*(w%4) .. 32*(w%4)+31).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM, same shape (used to re-zero an accumulator slot after it has been drained).
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace nesr
