// Thin inline-PTX wrappers for the sm_100a features the SDF kernels use:
// mbarrier, 1-D bulk async copy (TMA engine, SASS UBLKCP), tcgen05 (TMEM alloc,
// UMMA issue/commit, TMEM load) and the proxy / tcgen05 fences that order them.
// Nothing here is specific to the network; see chain.cuh / dw.cu for the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sdf {

// ----------------------------------------------------------------------------------------------
// Hang protection: every spin on an mbarrier is bounded.  A wrong phase/count in a kernel under
// development must abort the launch (trap -> cudaErrorLaunchFailure) instead of hanging the box.
// ----------------------------------------------------------------------------------------------
#ifndef SDF_SPIN_LIMIT
#define SDF_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Returns when the phase with the given parity has completed.  Traps after SDF_SPIN_LIMIT polls.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SDF_SPIN_LIMIT) {
      printf("sdf: mbarrier timeout block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Bulk async copies (non-tensor TMA).  Sizes and addresses are multiples of 16 bytes.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy encoding)
__device__ __forceinline__ void bulk_s2g_hint(void* gmem_dst, uint32_t smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_u32(void* gmem_dst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes) : "memory");
}
// writes of completed bulk stores (async proxy) -> visible to generic-proxy loads of threads that synchronise with this one
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// generic-proxy writes to shared memory -> visible to the async proxy (UMMA operand reads, bulk stores)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: descriptors.
//
// Shared-memory matrix descriptor (64 bit), field positions as in CUTLASS cute/arch/mma_sm100_desc.hpp
// (UMMA::SmemDescriptor): start_address[0,14) = addr>>4, leading_byte_offset[16,30) = LBO>>4,
// stride_byte_offset[32,46) = SBO>>4, version[46,48) = 1, base_offset[49,52) = 0, lbo_mode[52] = 0,
// layout_type[61,64) (0 = no swizzle / "interleave").
//
// All operands in this library use the NO-SWIZZLE canonical layouts:
//   K-major  (cute: ((8,m),(T,2)):((1T,SBO),(1,LBO)), T = 8 halves = 16 B):
//       8 rows x 16 B "core matrices" stored contiguously (128 B); SBO = distance between consecutive
//       8-row groups, LBO = distance between the two 8-element K granules of one K=16 instruction.
//   MN-major (cute: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO))):
//       core matrix = 8 K-rows x 16 B (8 contiguous MN elements); SBO = distance between consecutive
//       8-element MN blocks, LBO = distance between consecutive 8-row K groups.
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes, uint32_t layout_type = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}

// Instruction descriptor for tcgen05.mma.kind::f16 (UMMA::InstrDescriptor): c_format[4,6) (1 = F32),
// a_format[7,10), b_format[10,13) (0 = F16, 1 = BF16), a_major[15], b_major[16] (0 = K, 1 = MN),
// n_dim[17,23) = N>>3, m_dim[24,29) = M>>4.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, bool bf16, bool a_mn_major,
                                                            bool b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (bf16 ? 1u : 0u) << 7;
  d |= (bf16 ? 1u : 0u) << 10;
  d |= (a_mn_major ? 1u : 0u) << 15;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= static_cast<uint32_t>(N >> 3) << 17;
  d |= static_cast<uint32_t>(M >> 4) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the whole CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.ld: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32).  32x32b: thread i of the warp reads
// lane base+i, registers j = consecutive columns.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.st 32x32b.x16: thread i of the warp writes lane base+i, registers j -> consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One lane of the (fully active) warp.  Single-thread roles (TMA producer, MMA issuer) keep their loops warp-uniform
// and predicate only the issuing instructions with this: the loop scalars then live in uniform registers and a
// tcgen05.mma costs one instruction instead of ~25 (measured with tools/ring_probe.cu: 234 -> 138 cycles per
// M128 N256 K16 MMA through a 3 x 32 KB ring).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Warp-uniform wait: lane 0 polls (predicated try_wait, no divergent region), the warp votes on the outcome.
__device__ __forceinline__ void mbar_wait_vote(uint64_t* bar, uint32_t parity) {
  const bool poller = (threadIdx.x & 31) == 0;
  uint32_t spins = 0;
  while (true) {
    uint32_t ok = 0;
    if (poller) ok = mbar_try_wait(bar, parity) ? 1u : 0u;
    if (__any_sync(0xffffffffu, ok != 0)) break;
    if (++spins > SDF_SPIN_LIMIT) __trap();
  }
}

// the same on a shared-memory address (issue loops keep barrier addresses as integers)
__device__ __forceinline__ void mbar_wait_vote_u32(uint32_t bar_addr, uint32_t parity, bool poller) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok = 0;
    if (poller) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar_addr), "r"(parity)
          : "memory");
    }
    if (__any_sync(0xffffffffu, ok != 0)) break;
    if (++spins > SDF_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): the leader CTA (rank 0) issues M = 256 MMAs that read A / B halves
// from both CTAs' shared memory (same CTA-relative offsets) and write both CTAs' TMEM.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs completed) on the mbarrier at this CTA-relative offset in every CTA of mask
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// arrive on the mbarrier that lives at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
#ifdef SDF_REMOTE_ARRIVE_RELEASE_CLUSTER
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#else
  // default semantics (.release.cta), the form CUTLASS's ClusterBarrier::arrive(cta_id) uses for its 2-SM hand-shakes:
  // the .release.cluster form costs a MEMBAR.ALL + ERRBAR per arrival (ncu), which is pure latency in a relay warp
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#endif
}
// Wait on a barrier the peer CTA of the pair arrives on.  SDF_CLUSTER_WAIT_CTA = 1: default (.acquire.cta) semantics, the
// form CUTLASS's ClusterBarrier::wait uses -- what the waiter consumes after these barriers (peer shared memory / TMEM) is
// read by the tensor core through the async proxy, ordered by the peer's fence.proxy.async + tcgen05 fences, not by this
// thread's generic loads; the .acquire.cluster form makes ptxas follow every successful wait with CCTL.IVALL, which
// invalidates the SM's whole L1 (bias rows, lin_L row of the epilogue warps) a dozen times per layer step.
#ifndef SDF_CLUSTER_WAIT_CTA
#define SDF_CLUSTER_WAIT_CTA 1
#endif
#if SDF_CLUSTER_WAIT_CTA
#define SDF_CLUSTER_TRY_WAIT "mbarrier.try_wait.parity.shared::cta.b64"
#else
#define SDF_CLUSTER_TRY_WAIT "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64"
#endif
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      SDF_CLUSTER_TRY_WAIT " p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait for a phase completed (also) by arrivals from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > SDF_SPIN_LIMIT) {
      printf("sdf: cluster mbarrier timeout block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar),
             parity);
      __trap();
    }
  }
}

// warp-uniform forms on shared-memory addresses, acquire at cluster scope (barriers the peer CTA arrives on)
__device__ __forceinline__ void mbar_wait_cluster_vote_u32(uint32_t bar_addr, uint32_t parity, bool poller) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok = 0;
    if (poller) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          SDF_CLUSTER_TRY_WAIT " p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar_addr), "r"(parity)
          : "memory");
    }
    if (__any_sync(0xffffffffu, ok != 0)) break;
    if (++spins > SDF_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void umma_commit2_u32(uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
               "h"(cta_mask)
               : "memory");
}

}  // namespace sdf
