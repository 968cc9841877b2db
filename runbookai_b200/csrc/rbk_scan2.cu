// rbk_scan2.cu — K1 for batches of more than 128 queries: the same fused scan as
// rbk_scan.cu, on CTA PAIRS (thread-block cluster of 2, tcgen05 cta_group::2).
//
// Why: at B >= 256 the 1-CTA kernel is bound by L2->SM bandwidth (~10 TB/s measured, ncu
// captures in profiles/): every 128-query CTA pulls the whole 256-row corpus tile plus its
// query slab through L2 for each tile (589 KB per 128 q x 256 rows).  A pair computes a
// 256-query x 256-row tile with ONE copy of the corpus tile split across the two SMs (each
// CTA stages 128 corpus rows and its own 128 queries; the UMMA reads both halves).
//
//   scan2_kernel<false>  "streamed":  queries and corpus both flow through the smem ring
//                         (6 stages x 32 KB, K=64, 128-byte swizzle).  Any dim.
//                         L2->SM bytes per 256 q x 256 rows: 786 KB.
//   scan2_kernel<true>   "resident":  dim <= 768: the CTA's 128 queries stay in smem for the
//                         whole kernel (<= 12 panels x 16 KB, 128-byte swizzle) and only the
//                         corpus streams (4 stages x 8 KB, K=32, 64-byte swizzle).
//                         L2->SM bytes per 256 q x 256 rows: 393 KB.
//
// Protocol (leader = even CTA of the pair):
//   * both CTAs' TMA loads complete on the LEADER's barriers (count 1: the leader's
//     arrive.expect_tx covers the bytes of both CTAs);
//   * only the leader issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts the
//     "slot free" and "accumulator ready" arrivals to the same barrier offset in BOTH CTAs;
//   * each CTA's epilogue drains its own TMEM (its 128 queries) and arrives remotely on the
//     leader's tmem_empty barrier (count 8 = 4 warps x 2 CTAs).
#include "rbk_epilogue.cuh"
#include "rbk_internal.h"
#include "rbk_ptx.cuh"

namespace rbk {

namespace {

constexpr int kHalfN = kBlockN / 2;                  // corpus rows staged per CTA
constexpr int kTmemCols = 512;
constexpr int kMaxStages2 = 6;
constexpr int kPanelBytes = kBlockM * kBlockK * 2;   // 16 KiB: 128 rows x 64 bf16, SWIZZLE_128B

// streamed variant
constexpr int kStagesS = 6;
constexpr int kStageSBytes = kPanelBytes + kHalfN * kBlockK * 2;   // 32 KiB: query slab + corpus half-slab
// resident variant
#ifndef RBK_RES_K
#define RBK_RES_K 64
#endif
constexpr int kKR = RBK_RES_K;                        // bf16 per corpus stage: 64 (128-B rows) or 32 (64-B rows)
constexpr int kStagesR = kKR == 64 ? 2 : 4;
constexpr int kStageRBytes = kHalfN * kKR * 2;        // 16 KiB / 8 KiB

struct SmemTail2 {
  float invc[2][kBlockN];
  unsigned long long full[kMaxStages2];
  unsigned long long empty[kMaxStages2];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  unsigned long long a_full;
  uint32_t tmem_base;
};

// K-major operand, 64-byte swizzle: rows of 32 bf16 (64 B), 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t make_sw64_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;   // SWIZZLE_64B
  return d;
}

template <bool kRes>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kScanThreads, 1)
scan2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
             const __grid_constant__ CUtensorMap tmap_pf, const ScanParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  if (kRes && pad != 0) __trap();   // the resident layout has no slack: the base must be 1024-aligned
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = smem_u32(smem);
  constexpr int kStages = kRes ? kStagesR : kStagesS;
  constexpr int kStageBytes = kRes ? kStageRBytes : kStageSBytes;
  // resident: [A panels: num_kb x 16 KB][corpus ring][tail]; streamed: [ring][tail]
  const uint32_t ring_off = kRes ? static_cast<uint32_t>(p.num_kb) * kPanelBytes : 0u;
  SmemTail2* tail = reinterpret_cast<SmemTail2*>(smem + ring_off + kStages * kStageBytes);
  const uint32_t ring_base = smem_base + ring_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int qb = pair % p.QB;                        // 256-query block
  const int r = pair / p.QB;                         // corpus range
  const int t0 = static_cast<int>(static_cast<long long>(p.n_tiles) * r / p.R);
  const int t1 = static_cast<int>(static_cast<long long>(p.n_tiles) * (r + 1) / p.R);
  const int q_row0 = qb * 2 * kBlockM + static_cast<int>(rank) * kBlockM;   // this CTA's first query
  // corpus k-steps per tile: streamed K=64 per stage, resident K=32 per stage
  const int n_ksteps = kRes ? (p.dpad + kKR - 1) / kKR : p.num_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);
      mbar_init(smem_u32(&tail->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tail->tmem_full[a]), 1);
      mbar_init(smem_u32(&tail->tmem_empty[a]), 8);
    }
    mbar_init(smem_u32(&tail->a_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(smem_u32(&tail->tmem_base), kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, elected issue; both CTAs) =====================
    if (kRes) {
      // this CTA's 128 queries, all K, once: 128-byte-swizzled panels of 64 columns
      const uint32_t a_full = smem_u32(&tail->a_full);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(a_full, 2u * static_cast<uint32_t>(p.num_kb) * kPanelBytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d_2cta(smem_base + kb * kPanelBytes, &tmap_q, a_full, kb * kBlockK, q_row0);
      }
      __syncwarp();
    }
    volatile int* prog = p.progress + r * p.QB;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      if (rank == 0 && lane == 0) lockstep_pace(prog, p.QB, qb, tile - t0);
      __syncwarp();
      const int c_row0 = tile * kBlockN + static_cast<int>(rank) * kHalfN;
      if (kRes && p.prefetch_tiles > 0 && tile + p.prefetch_tiles < t1 && lane == 0) {
        // the resident layout leaves only 32 KB of smem ring per CTA: fetch this CTA's rows of a
        // later tile into L2 now so the ring's loads are L2 hits (prefetch boxes: 128 rows x 256 cols)
        for (int c = 0; c < p.dpad; c += 256) tma_prefetch_l2_2d(&tmap_pf, c, c_row0 + p.prefetch_tiles * kBlockN);
      }
      for (int ks = 0; ks < n_ksteps; ++ks) {
        mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&tail->full[s]);
        const uint32_t dst = ring_base + s * kStageBytes;
        if (elect_one()) {
          // Only the leader arrives (once, expecting BOTH CTAs' bytes).  The peer's bytes may land
          // first and drive the tx-count negative; the phase cannot complete before the leader's
          // arrive, and the peer re-uses a slot only after the leader's MMAs consumed it.
          if (rank == 0) mbar_arrive_expect_tx(full, 2 * kStageBytes);
          if (kRes) {
            tma_load_2d_2cta(dst, &tmap_c, full, ks * kKR, c_row0);
          } else {
            tma_load_2d_2cta(dst, &tmap_q, full, ks * kBlockK, q_row0);
            tma_load_2d_2cta(dst + kPanelBytes, &tmap_c, full, ks * kBlockK, c_row0);
          }
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    if (rank == 0 && lane == 0 && p.QB > 1) prog[qb] = 0x7FFFFFFF;   // done: never hold a peer back
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp, elected issue) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, kBlockN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      if (kRes) {
        mbar_wait(smem_u32(&tail->a_full), 0u);
        tc_fence_after();
      }
      for (int tile = t0; tile < t1; ++tile) {
        mbar_wait(smem_u32(&tail->tmem_empty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kBlockN);
        for (int ks = 0; ks < n_ksteps; ++ks) {
          mbar_wait(smem_u32(&tail->full[s]), ph);
          tc_fence_after();
          const uint32_t st = ring_base + s * kStageBytes;
          if (elect_one()) {
            if (kRes) {
              if constexpr (kKR == 32) {
                // stage = 32 corpus columns; queries: panel ks/2, 64-byte half (ks&1) of its 128-byte rows
                const uint32_t a0 = smem_base + (ks >> 1) * kPanelBytes + (ks & 1) * 64;
#pragma unroll
                for (int k = 0; k < kKR / 16; ++k)
                  umma_bf16_ss_2cta(d_tmem, make_sw128_kmajor_desc(a0 + k * 32), make_sw64_kmajor_desc(st + k * 32),
                                    idesc, (ks | k) != 0 ? 1u : 0u);
              } else {
                const uint32_t a0 = smem_base + ks * kPanelBytes;
#pragma unroll
                for (int k = 0; k < kKR / 16; ++k)
                  umma_bf16_ss_2cta(d_tmem, make_sw128_kmajor_desc(a0 + k * 32),
                                    make_sw128_kmajor_desc(st + k * 32), idesc, (ks | k) != 0 ? 1u : 0u);
              }
            } else {
              const uint64_t adesc0 = make_sw128_kmajor_desc(st);
              const uint64_t bdesc0 = make_sw128_kmajor_desc(st + kPanelBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)   // +32 bytes per k-step = +2 in the 16-byte address field
                umma_bf16_ss_2cta(d_tmem, adesc0 + static_cast<uint64_t>(2 * k), bdesc0 + static_cast<uint64_t>(2 * k),
                                  idesc, (ks | k) != 0 ? 1u : 0u);
            }
            umma_commit_2cta(smem_u32(&tail->empty[s]));
            if (ks == n_ksteps - 1) umma_commit_2cta(smem_u32(&tail->tmem_full[as]));
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: thread <-> query (both CTAs) =====================
    run_epilogue<true>(p, tail->invc, tail->tmem_full, tail->tmem_empty, tmem_base, qb, r, rank, t0, t1, warp, lane);
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's smem/TMEM must outlive the leader's last MMA and all remote arrives
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

constexpr size_t kMaxSmem = 232448;   // 227 KiB opt-in limit per CTA on sm_100

}  // namespace

// Can the query block stay resident for this padded dim?
int scan2_resident_k() { return kKR; }

bool scan2_resident_fits(int dpad) {
  const size_t num_kb = static_cast<size_t>((dpad + kBlockK - 1) / kBlockK);
  return num_kb * kPanelBytes + static_cast<size_t>(kStagesR) * kStageRBytes + sizeof(SmemTail2) <= kMaxSmem;
}

// p.QB counts 256-query blocks, p.R CTA pairs per block; grid = 2 * QB * R CTAs.
// streamed: tmap_c has 128-row x 64-col boxes (SWIZZLE_128B); resident: the same, or 128-row x 32-col
// (SWIZZLE_64B) when built with RBK_RES_K=32.
cudaError_t launch_scan2(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const CUtensorMap& tmap_pf,
                         const ScanParams& p, bool resident, cudaStream_t stream) {
  cudaError_t e;
  if (resident) {
    const size_t smem = static_cast<size_t>(p.num_kb) * kPanelBytes + static_cast<size_t>(kStagesR) * kStageRBytes +
                        sizeof(SmemTail2);
    e = cudaFuncSetAttribute(scan2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    scan2_kernel<true><<<2 * p.QB * p.R, kScanThreads, smem, stream>>>(tmap_q, tmap_c, tmap_pf, p);
  } else {
    const size_t smem = static_cast<size_t>(kStagesS) * kStageSBytes + sizeof(SmemTail2) + 1024;
    e = cudaFuncSetAttribute(scan2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    scan2_kernel<false><<<2 * p.QB * p.R, kScanThreads, smem, stream>>>(tmap_q, tmap_c, tmap_pf, p);
  }
  return cudaGetLastError();
}

}  // namespace rbk
