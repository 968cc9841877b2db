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
//                         (7 stages x 32 KB, K=64, 128-byte swizzle).  Any dim.
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
constexpr int kMaxStages2 = 7;
constexpr int kPanelBytes = kBlockM * kBlockK * 2;   // 16 KiB: 128 rows x 64 bf16, SWIZZLE_128B

// streamed variant
#ifndef RBK_STAGES_S
#define RBK_STAGES_S 7
#endif
constexpr int kStagesS = RBK_STAGES_S;
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

template <bool kRes, int kHalves>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * kHalves, 1)
scan2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
             const __grid_constant__ CUtensorMap tmap_pf, const ScanParams p, const int n_stages_s) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  {
    // the launch adds 1 KB of alignment slack only when it fits under the 227 KB limit; without slack the
    // (in practice always 1024-aligned) dynamic smem base must really be aligned
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    const uint32_t need = (kRes ? static_cast<uint32_t>(p.num_kb) * kPanelBytes + kStagesR * kStageRBytes
                                : static_cast<uint32_t>(n_stages_s) * kStageSBytes) + static_cast<uint32_t>(sizeof(SmemTail2));
    if (pad + need > dyn) __trap();
  }
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = smem_u32(smem);
  const int kStages = kRes ? kStagesR : n_stages_s;   // streamed: 7 when the smem base is 1024-aligned, else 6 + slack
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
  const int t0 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * r / p.R_local);
  const int t1 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * (r + 1) / p.R_local);
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
      mbar_init(smem_u32(&tail->tmem_empty[a]), 8 * kHalves);   // one arrive per epilogue warp of both CTAs
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
    volatile int* prog = p.progress + p.prog_base + r * p.QB;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      if (rank == 0 && lane == 0) lockstep_pace(prog, p.QB, qb, tile - t0, p.max_lead_tiles);
      __syncwarp();
      const int c_row0 = tile * kBlockN + static_cast<int>(rank) * kHalfN;
      if (p.prefetch_tiles > 0 && tile + p.prefetch_tiles < t1 && lane == 0) {
        // HBM latency outside the smem ring: fetch this CTA's rows of a later tile into L2 now so that the ring's
        // loads are L2 hits (prefetch boxes: 128 rows x 256 cols).  The ring holds 7 x 16 KB of corpus per SM
        // (16.6 MB chip-wide), which at ~2 us of HBM latency caps a pure stream near 5.5 TB/s of the 7.3 TB/s that
        // TMA reads reach on this chip (scripts/probes/tma_stream_probe.cu).
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
    run_epilogue<true, kHalves>(p, tail->invc, tail->tmem_full, tail->tmem_empty, tmem_base, qb, p.unit_base + r, rank,
                                t0, t1, warp, lane);
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's smem/TMEM must outlive the leader's last MMA and all remote arrives
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

constexpr size_t kMaxSmem = 232448;   // 227 KiB opt-in limit per CTA on sm_100

__global__ void smem_align_probe_kernel(unsigned int* out) {
  extern __shared__ __align__(1024) uint8_t probe_smem[];
  if (threadIdx.x == 0) *out = smem_u32(probe_smem) & 1023u;
}

// Is the dynamic shared memory window 1024-byte aligned for a near-maximal allocation?  (It is on every
// driver seen so far; the answer is cached per device.)
bool smem_base_is_aligned() {
  static int cached[64];   // 0 = unknown, 1 = aligned, 2 = not
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (cached[dev] == 0) {
    unsigned int* d = nullptr;
    unsigned int h = 1;
    const int smem = 200 * 1024;
    if (cudaMalloc(reinterpret_cast<void**>(&d), 4) == cudaSuccess &&
        cudaFuncSetAttribute(smem_align_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) ==
            cudaSuccess) {
      smem_align_probe_kernel<<<1, 32, smem>>>(d);
      if (cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost) != cudaSuccess) h = 1;
    }
    cudaFree(d);
    cached[dev] = h == 0 ? 1 : 2;
  }
  return cached[dev] == 1;
}

#ifdef RBK_EXPERIMENTAL
// ---------------------------------------------------------------------------------------------
// Hybrid variant: the first `res_kb` 64-column panels of the CTA's queries stay resident in smem,
// the rest of the query slab and the whole corpus stream through a ring of 16 KB SLOTS (a k-block
// whose query panel is resident takes one slot, a fully streamed k-block two).  Compared with the
// fully streamed kernel this removes res_kb/num_kb of the query bytes from the L2->SM stream
// (d=768, res_kb=6: 288 KB instead of 384 KB per CTA per 256x256 tile) while keeping a 128 KB
// ring — the fully resident layout (32 KB ring) was latency-bound.
constexpr int kSlotBytes = kPanelBytes;   // 16 KiB: 128 rows x 64 bf16, SWIZZLE_128B (queries or corpus half)
constexpr int kMaxSlots = 10;

struct SmemTailH {
  float invc[2][kBlockN];
  unsigned long long full[kMaxSlots];
  unsigned long long empty[kMaxSlots];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  unsigned long long a_full;
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kScanThreads, 1)
scan2h_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
              const ScanParams p, const int res_kb, const int n_slots) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  if (pad != 0) __trap();   // the layout is sized without slack: the dynamic smem base must be 1024-aligned
  uint8_t* smem = smem_raw;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring_base = smem_base + static_cast<uint32_t>(res_kb) * kPanelBytes;
  SmemTailH* tail = reinterpret_cast<SmemTailH*>(smem + (res_kb + n_slots) * kPanelBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int qb = pair % p.QB;
  const int r = pair / p.QB;
  const int t0 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * r / p.R_local);
  const int t1 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * (r + 1) / p.R_local);
  const int q_row0 = qb * 2 * kBlockM + static_cast<int>(rank) * kBlockM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < n_slots; ++s) {
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, elected issue; both CTAs) =====================
    if (res_kb > 0) {
      const uint32_t a_full = smem_u32(&tail->a_full);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(a_full, 2u * static_cast<uint32_t>(res_kb) * kPanelBytes);
        for (int kb = 0; kb < res_kb; ++kb)
          tma_load_2d_2cta(smem_base + kb * kPanelBytes, &tmap_q, a_full, kb * kBlockK, q_row0);
      }
      __syncwarp();
    }
    volatile int* prog = p.progress + p.prog_base + r * p.QB;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      if (rank == 0 && lane == 0) lockstep_pace(prog, p.QB, qb, tile - t0, p.max_lead_tiles);
      __syncwarp();
      const int c_row0 = tile * kBlockN + static_cast<int>(rank) * kHalfN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int n_loads = kb < res_kb ? 1 : 2;   // [query panel,] corpus half-slab
        for (int l = 2 - n_loads; l < 2; ++l) {
          mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
          const uint32_t full = smem_u32(&tail->full[s]);
          const uint32_t dst = ring_base + s * kSlotBytes;
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(full, 2 * kSlotBytes);   // both CTAs' bytes (see scan2_kernel)
            if (l == 0) tma_load_2d_2cta(dst, &tmap_q, full, kb * kBlockK, q_row0);
            else tma_load_2d_2cta(dst, &tmap_c, full, kb * kBlockK, c_row0);
          }
          __syncwarp();
          if (++s == n_slots) { s = 0; ph ^= 1u; }
        }
      }
    }
    if (rank == 0 && lane == 0 && p.QB > 1) prog[qb] = 0x7FFFFFFF;
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp, elected issue) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, kBlockN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      if (res_kb > 0) {
        mbar_wait(smem_u32(&tail->a_full), 0u);
        tc_fence_after();
      }
      for (int tile = t0; tile < t1; ++tile) {
        mbar_wait(smem_u32(&tail->tmem_empty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kBlockN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const bool streamed_a = kb >= res_kb;
          uint32_t a_addr = smem_base + kb * kPanelBytes;
          int sa = -1;
          if (streamed_a) {
            mbar_wait(smem_u32(&tail->full[s]), ph);
            a_addr = ring_base + s * kSlotBytes;
            sa = s;
            if (++s == n_slots) { s = 0; ph ^= 1u; }
          }
          mbar_wait(smem_u32(&tail->full[s]), ph);
          tc_fence_after();
          const uint64_t adesc0 = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc0 = make_sw128_kmajor_desc(ring_base + s * kSlotBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_ss_2cta(d_tmem, adesc0 + static_cast<uint64_t>(2 * k), bdesc0 + static_cast<uint64_t>(2 * k),
                                idesc, (kb | k) != 0 ? 1u : 0u);
            if (sa >= 0) umma_commit_2cta(smem_u32(&tail->empty[sa]));
            umma_commit_2cta(smem_u32(&tail->empty[s]));
            if (kb == p.num_kb - 1) umma_commit_2cta(smem_u32(&tail->tmem_full[as]));
          }
          __syncwarp();
          if (++s == n_slots) { s = 0; ph ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
  } else {
    run_epilogue<true>(p, tail->invc, tail->tmem_full, tail->tmem_empty, tmem_base, qb, p.unit_base + r, rank, t0, t1, warp,
                       lane);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

#endif  // RBK_EXPERIMENTAL

}  // namespace

// (exported for the cluster kernel's launch, rbk_scan4.cu)
bool scan_smem_base_is_aligned() { return smem_base_is_aligned(); }

// Can the query block stay resident for this padded dim?
int scan2_resident_k() { return kKR; }

bool scan2_resident_fits(int dpad) {
  const size_t num_kb = static_cast<size_t>((dpad + kBlockK - 1) / kBlockK);
  return num_kb * kPanelBytes + static_cast<size_t>(kStagesR) * kStageRBytes + sizeof(SmemTail2) <= kMaxSmem;
}

// p.QB counts 256-query blocks, p.R CTA pairs per block; grid = 2 * QB * R CTAs.
// streamed: tmap_c has 128-row x 64-col boxes (SWIZZLE_128B); resident: the same, or 128-row x 32-col
// (SWIZZLE_64B) when built with RBK_RES_K=32.
template <bool kRes, int kHalves>
static cudaError_t launch_scan2_t(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const CUtensorMap& tmap_pf,
                                  const ScanParams& p, size_t smem, int n_stages, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(scan2_kernel<kRes, kHalves>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  scan2_kernel<kRes, kHalves><<<2 * p.QB * p.R_local, 64 + 128 * kHalves, smem, stream>>>(tmap_q, tmap_c, tmap_pf, p, n_stages);
  return cudaGetLastError();
}

// halves: 1 = four epilogue warps per CTA, 2 = eight (two lists per unit and query; see run_epilogue).
cudaError_t launch_scan2(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const CUtensorMap& tmap_pf,
                         const ScanParams& p, bool resident, int halves, cudaStream_t stream, int* ring_stages_out) {
#ifdef RBK_EXPERIMENTAL
  if (resident) {
    const size_t smem = static_cast<size_t>(p.num_kb) * kPanelBytes + static_cast<size_t>(kStagesR) * kStageRBytes +
                        sizeof(SmemTail2);
    return halves == 2 ? launch_scan2_t<true, 2>(tmap_q, tmap_c, tmap_pf, p, smem, kStagesS, stream)
                       : launch_scan2_t<true, 1>(tmap_q, tmap_c, tmap_pf, p, smem, kStagesS, stream);
  }
#else
  if (resident) return cudaErrorNotSupported;
#endif
  // 7 stages fill the 227 KB exactly (no alignment slack): only when the dynamic smem base is 1024-aligned on
  // this device/driver (probed once); otherwise 6 stages + 1 KB of slack
  const int n_stages = smem_base_is_aligned() ? kStagesS : kStagesS - 1;
  size_t smem = static_cast<size_t>(n_stages) * kStageSBytes + sizeof(SmemTail2);
  if (smem + 1024 <= kMaxSmem) smem += 1024;
  if (ring_stages_out) *ring_stages_out = n_stages;
#ifdef RBK_EXPERIMENTAL
  if (halves != 2) return launch_scan2_t<false, 1>(tmap_q, tmap_c, tmap_pf, p, smem, n_stages, stream);
#else
  if (halves != 2) return cudaErrorNotSupported;   // the shipped kernel always runs eight epilogue warps
#endif
  return launch_scan2_t<false, 2>(tmap_q, tmap_c, tmap_pf, p, smem, n_stages, stream);
}

#ifdef RBK_EXPERIMENTAL
// Hybrid launch: res_kb resident query panels + n_slots ring slots (16 KB each) must fit 227 KB.
cudaError_t launch_scan2h(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c_half, const ScanParams& p, int res_kb,
                          int n_slots, cudaStream_t stream) {
  if (res_kb > p.num_kb) res_kb = p.num_kb;
  if (n_slots > kMaxSlots) n_slots = kMaxSlots;
  if (n_slots < 4) n_slots = 4;
  while (static_cast<size_t>(res_kb + n_slots) * kPanelBytes + sizeof(SmemTailH) > kMaxSmem && res_kb > 0) --res_kb;
  const size_t smem = static_cast<size_t>(res_kb + n_slots) * kPanelBytes + sizeof(SmemTailH);
  cudaError_t e = cudaFuncSetAttribute(scan2h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan2h_kernel<<<2 * p.QB * p.R_local, kScanThreads, smem, stream>>>(tmap_q, tmap_c_half, p, res_kb, n_slots);
  return cudaGetLastError();
}

#endif  // RBK_EXPERIMENTAL

}  // namespace rbk
