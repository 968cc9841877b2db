// rbk_ptx.cuh — thin inline-PTX wrappers for the sm_100a features the scan kernel uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// shared-memory matrix descriptor.  Hand-written for this engine; bit layouts follow the
// PTX ISA tables for tcgen05 descriptors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rbk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure
// -> RBK_ECUDA), never as a hung GPU.  ~2^31 cycles is seconds; real waits are micro-
// seconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3FFu) == 0 && (clock64() - t0) > (1ll << 31)) __trap();
  }
}

// ---------------------------------------------------------------- TMA
// Pull one 128-byte line into L1 (no register result, no scoreboard wait).
__device__ __forceinline__ void prefetch_l1(const void* gptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<uint64_t>(gptr)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes).
// crd0 = innermost (element) coordinate, crd1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t crd0,
                                            int32_t crd1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(crd0), "r"(crd1)
      : "memory");
}

// Pull a box into L2 only (no smem, no barrier): hides HBM latency when the smem ring is shallow.
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int32_t crd0, int32_t crd1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(crd0), "r"(crd1)
               : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread are done
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// Warp-collective: lane t reads TMEM lane (taddr.lane + t), 32 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, with the loaded registers as in/out operands: a data dependency that keeps every use of
// v[] after the wait even when the load was issued several statements earlier (software pipelining).
__device__ __forceinline__ void tmem_wait_ld_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 64 bf16
// (128 B) packed densely: 8-row groups are 1024 B apart (SBO); LBO is unused for
// swizzled K-major layouts (encoded 1); version field = 1 (sm_100); layout = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                      // [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                      // [46,48) descriptor version
  d |= static_cast<uint64_t>(2) << 61;                      // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense, M x N.
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int M, int N) {
  return (1u << 4)                              // [4,6)   D format  = F32
         | (1u << 7)                            // [7,10)  A format  = BF16
         | (1u << 10)                           // [10,13) B format  = BF16
         | (static_cast<uint32_t>(N >> 3) << 17)  // [17,23) N >> 3
         | (static_cast<uint32_t>(M >> 4) << 24); // [24,29) M >> 4
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// In the shared::cluster window the CTA rank sits above bit 24; clearing bit 24 of a local
// address names the same offset in the even (leader) CTA of the pair.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same offset in the leader CTA (works from either CTA).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {
  // default semantics (release at CTA scope): a cluster-scope release would drag in a
  // MEMBAR.GPU + ERRBAR per arrive (seen in the ncu source view) for no benefit here
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(local_bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of a pair into ITS OWN smem; the bytes are accounted on the
// LEADER's mbarrier (the MMA issuer waits there for both halves).
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const void* tmap, uint32_t local_bar,
                                                 int32_t crd0, int32_t crd1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(local_bar & kPeerBitMask), "r"(crd0), "r"(crd1)
      : "memory");
}
// The same load, MULTICAST: the box is written at the same smem offset of every CTA in cta_mask (bit i = cluster
// rank i) with ONE read of L2, and each destination's bytes are accounted on the barrier at this offset in the
// leader (even rank) of THAT destination's pair.
__device__ __forceinline__ void tma_load_2d_2cta_mc(uint32_t smem_dst, const void* tmap, uint32_t local_bar,
                                                    int32_t crd0, int32_t crd1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(local_bar & kPeerBitMask), "h"(cta_mask),
        "r"(crd0), "r"(crd1)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T; M = 256.
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once the pair's MMAs retire) on the barrier at this offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// Same, to the CTAs named by cta_mask (bit i = cluster rank i): clusters of more than one pair.
__device__ __forceinline__ void umma_commit_2cta_mask(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(bar), "h"(cta_mask)
      : "memory");
}

// One lane of a converged warp.  The producer and MMA roles run their loops with the WHOLE warp
// (uniform control flow, operands in uniform registers) and elect a lane only around the
// asynchronous instruction itself: a loop that lives inside `if (lane == 0)` makes every operand
// a per-thread value and costs an ELECT + R2UR round trip per descriptor (~16 instructions per
// tcgen05.mma, seen in the ncu source view), which made 32-cycle N=64 MMAs issue-bound.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Order-preserving map float -> u32 (larger float -> larger integer; NaN never packed).
__device__ __forceinline__ uint32_t f32_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_ordered(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}
// Candidate key: (score desc, row asc) == larger key first.
__device__ __forceinline__ uint64_t pack_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_ordered(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0xFFFFFFFFu - static_cast<uint32_t>(k); }
__device__ __forceinline__ float key_score(uint64_t k) { return f32_from_ordered(static_cast<uint32_t>(k >> 32)); }

}  // namespace rbk
