// Inline-PTX wrappers for the sm_100a features the score kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma kind::i8 / commit / ld / fences).
// Written against the PTX ISA as exercised by the CUTLASS/CuTe sm100 headers; no CUTLASS code
// is included or linked.
#pragma once
#include <cstdint>
#include <cuda.h>

namespace smb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Non-blocking phase test.  Issued well before its result is needed, its ~90 clk latency (the same an already
// complete try_wait costs) hides behind whatever the thread does in between.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// try_wait that lets the hardware keep the thread asleep for up to `ns` before reporting failure: the
// waiting warp stops competing for issue slots / the ALU pipe with the warps doing the epilogue arithmetic.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must trap (launch error, reported through the C ABI), never hang the box.
#ifndef SMB_MBAR_SPIN_LIMIT
#define SMB_MBAR_SPIN_LIMIT (1u << 22)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
#ifdef SMB_MBAR_BUSY_POLL
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (SMB_MBAR_SPIN_LIMIT << 4)) __trap();
  }
#else
  while (!mbar_try_wait_hint(bar, parity, 4000u)) {
    if (++spins > SMB_MBAR_SPIN_LIMIT) __trap();
  }
#endif
}

// ---------------------------------------------------------------- register reallocation between warpgroups
// (all four warps of a warpgroup must execute the same instruction; N a multiple of 8)
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 1-D bulk copy global -> shared (no tensor map), completion counted in bytes on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_result_addr) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, u8 x u8 -> s32, issued by ONE thread.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 64 consecutive columns in ONE instruction, delivered into two 32-register runs.  Two of these per
// warp and tile instead of four .x32: ptxas keeps them back to back (it started the first run's max tree before
// the fourth .x32 load), so the TMEM buffer is handed back ~20 ALU instructions earlier.
__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t (&v0)[32], uint32_t (&v1)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v0[8]), "=r"(v0[9]), "=r"(v0[10]), "=r"(v0[11]), "=r"(v0[12]), "=r"(v0[13]), "=r"(v0[14]), "=r"(v0[15]), "=r"(v0[16]), "=r"(v0[17]), "=r"(v0[18]), "=r"(v0[19]), "=r"(v0[20]), "=r"(v0[21]), "=r"(v0[22]), "=r"(v0[23]), "=r"(v0[24]), "=r"(v0[25]), "=r"(v0[26]), "=r"(v0[27]), "=r"(v0[28]), "=r"(v0[29]), "=r"(v0[30]), "=r"(v0[31]), "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7]), "=r"(v1[8]), "=r"(v1[9]), "=r"(v1[10]), "=r"(v1[11]), "=r"(v1[12]), "=r"(v1[13]), "=r"(v1[14]), "=r"(v1[15]), "=r"(v1[16]), "=r"(v1[17]), "=r"(v1[18]), "=r"(v1[19]), "=r"(v1[20]), "=r"(v1[21]), "=r"(v1[22]), "=r"(v1[23]), "=r"(v1[24]), "=r"(v1[25]), "=r"(v1[26]), "=r"(v1[27]), "=r"(v1[28]), "=r"(v1[29]), "=r"(v1[30]), "=r"(v1[31])
               : "r"(taddr)
               : "memory");
}

// ---------------------------------------------------------------- CTA pair (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// In the shared::cluster window bit 24 selects the peer CTA of a pair; clearing it addresses CTA 0's copy.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// 2-D tiled load into THIS CTA's shared memory whose byte count is reported to the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int32_t c0,
                                                int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_512_2sm(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_result_addr) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T over the CTA pair: M = 256 (128 rows per CTA), each CTA supplies its own
// A rows and half of the N columns of B.  Issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all previously issued MMAs completed) on the barrier at this offset in every CTA of cta_mask.
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 B (the whole K
// extent of a SIFT descriptor), 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled
// K-major layouts (encoded 1 like CuTe does); descriptor version 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;            // version
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::i8: D = s32, A = B = unsigned 8-bit, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc_u8u8s32(uint32_t m, uint32_t n) {
  return (2u << 4)            // c_format: S32
         | (0u << 7)          // a_format: UINT8
         | (0u << 10)         // b_format: UINT8
         | (0u << 15)         // a_major: K
         | (0u << 16)         // b_major: K
         | ((n >> 3) << 17)   // n_dim
         | ((m >> 4) << 24);  // m_dim
}

}  // namespace ptx
}  // namespace smb
