// rbk_scan4.cu — K1 for batches of more than 128 queries on CLUSTERS OF TWO CTA PAIRS (4 CTAs) with one
// operand of every k-block delivered by TMA MULTICAST.
//
// Why it was built, and what it measured (round 2).  Every pair of the pair kernel (rbk_scan2.cu) pulls
// (256 + 256) x 768 x 2 B = 786 KB through the L2->SM fabric per 256 x 256 tile and its tensor pipe is busy 82 % of
// the SM's active cycles (profiles/r01_scan_cfg3_final2.txt).  Two pairs that need the SAME operand slab can have it
// read from L2 once and written into both pairs' shared memory by one multicast TMA: 590 KB of L2 reads per
// pair-tile instead of 786 KB.  Result on B200 (profiles/r02_scan4_cfg3_128sm.txt): L2 slice traffic -22 %, tensor
// pipe 96.7 % of active cycles - but only 33 clusters of 4 CTAs fit the chip (GPCs whose SM count is not a
// multiple of 4 leave SMs over; scripts/probes/cluster_probe.cu), i.e. 128-132 of 148 SMs, and under the 1 kW power
// cap every variant - pair kernel on 144 SMs, this kernel on 128, this kernel plus a concurrent pair-kernel tail on
// the spare SMs (rbk_capi.cu, run_scan) - lands within 2 % of the others (profiles/r02_ab_cluster_multicast_*.jsonl:
// cfg3 10.95-11.15 ms, cfg2 0.286 ms, cfg5 1.49-1.51 ms, cfg4 shard 4.35-4.45 ms on one box).  The long scans are
// bound by power, not by the fabric, so the simpler kernel ships and this one stays behind -DRBK_EXPERIMENTAL.
//
// Two sharing modes, chosen per launch:
//   kShareC  (even number of 256-query blocks): the pairs of a cluster hold DIFFERENT query blocks and walk the
//            same corpus tiles; every corpus half-slab (128 rows x 64 cols per CTA rank parity) is fetched as two
//            64-row boxes, one by each pair, multicast to the same-parity CTA of both pairs.
//   !kShareC (odd number of query blocks, e.g. B = 256): the pairs hold the SAME query block and take alternate
//            tiles of the cluster's corpus range; the query slab is what is multicast.
//
// Protocol (cluster ranks 0..3; pair = rank >> 1; c = rank & 1; pair leader = its even rank):
//   * full[s]  (leader of each pair, count 1): the leader's arrive.expect_tx covers all 64 KB that land in ITS pair's
//     two CTAs for the stage - 32 KB private operand (own loads) + 32 KB shared operand (two 8 KB boxes into each
//     CTA, one issued by each pair; a multicast box signals, in every destination, the barrier of that
//     destination's pair leader);
//   * empty[s] (every CTA, count 2): a slot is written by BOTH pairs' producers (the shared boxes), so it is free
//     only when both pairs' MMAs have consumed it: each leader's tcgen05.commit is multicast to all four CTAs;
//   * tmem_full / tmem_empty stay inside a pair (commit mask = the pair's two ranks), as in rbk_scan2.cu.
// The two pairs therefore advance in lockstep at smem-stage granularity, by construction.
#ifdef RBK_EXPERIMENTAL   // measured (profiles/r02_scan4_cfg3_128sm.txt, r02_ab_cluster_multicast_*.jsonl): not in the default build
#include "rbk_epilogue.cuh"
#include "rbk_internal.h"
#include "rbk_ptx.cuh"

namespace rbk {

namespace {

constexpr int kHalfN = kBlockN / 2;                  // corpus rows staged per CTA
constexpr int kTmemCols = 512;
constexpr int kMaxStages4 = 7;
constexpr int kPanelBytes = kBlockM * kBlockK * 2;   // 16 KiB: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kBoxBytes = kPanelBytes / 2;           // 8 KiB: one 64-row multicast box
constexpr int kStageBytes = 2 * kPanelBytes;         // 32 KiB: query slab + corpus half-slab

struct SmemTail4 {
  float invc[2][kBlockN];
  unsigned long long full[kMaxStages4];
  unsigned long long empty[kMaxStages4];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  uint32_t tmem_base;
};

template <bool kShareC>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(64 + 128 * 2, 1)
scan4_kernel(const __grid_constant__ CUtensorMap tmap_q128, const __grid_constant__ CUtensorMap tmap_q64,
             const __grid_constant__ CUtensorMap tmap_c128, const __grid_constant__ CUtensorMap tmap_c64,
             const ScanParams p, const int n_stages) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + static_cast<uint32_t>(n_stages) * kStageBytes + static_cast<uint32_t>(sizeof(SmemTail4)) > dyn) __trap();
  }
  uint8_t* smem = smem_raw + pad;
  const uint32_t ring_base = smem_u32(smem);
  SmemTail4* tail = reinterpret_cast<SmemTail4*>(smem + n_stages * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0..3
  const int pr = static_cast<int>(rank >> 1);        // pair inside the cluster
  const int c = static_cast<int>(rank & 1u);         // CTA inside the pair (0 = leader)
  const int cl = blockIdx.x >> 2;                    // cluster index
  // columns = clusters that share a corpus range (lockstep-paced through p.progress)
  const int n_cols = kShareC ? p.QB / 2 : p.QB;
  const int col = cl % n_cols;
  const int rc = cl / n_cols;                        // corpus range of this cluster
  const int qb = kShareC ? 2 * col + pr : col;       // 256-query block of this pair
  const int t0 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * rc / p.RC);
  const int t1 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * (rc + 1) / p.RC);
  // this pair's tiles: t_first, t_first + t_step, ... (n_iter of them; indices >= t1 are phantoms)
  const int t_step = kShareC ? 1 : 2;
  const int t_first = kShareC ? t0 : t0 + pr;
  const int n_iter = kShareC ? t1 - t0 : (t1 - t0 + 1) / 2;
  const int unit = p.unit_base + (kShareC ? rc : 2 * rc + pr);       // list / publisher index inside the query block (p.R of them)
  const uint16_t mc_mask = static_cast<uint16_t>(0x5u << c);              // same-parity CTA of both pairs
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * pr));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q128);
    tma_prefetch_desc(&tmap_q64);
    tma_prefetch_desc(&tmap_c128);
    tma_prefetch_desc(&tmap_c64);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);
      mbar_init(smem_u32(&tail->empty[s]), 2);        // one multicast commit from each pair's leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tail->tmem_full[a]), 1);
      mbar_init(smem_u32(&tail->tmem_empty[a]), 16);   // 8 epilogue warps x 2 CTAs of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(smem_u32(&tail->tmem_base), kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of ALL FOUR CTAs initialised before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, elected issue; all four CTAs) =====================
    volatile int* prog = p.progress + p.prog_base + rc * n_cols;
    const int q_row0 = qb * 2 * kBlockM + c * kBlockM;   // this CTA's first query
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_iter; ++it) {
      int tile = t_first + it * t_step;
      if (tile >= t1) tile = p.n_tiles;                  // phantom: rows beyond n_rows, TMA zero-fills
      if (rank == 0 && lane == 0) lockstep_pace(prog, n_cols, col, it, kShareC ? p.max_lead_tiles : (p.max_lead_tiles + 1) / 2);
      __syncwarp();
      const int c_row0 = tile * kBlockN + c * kHalfN;
      for (int ks = 0; ks < p.num_kb; ++ks) {
        mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);   // both pairs have consumed this slot
        const uint32_t full = smem_u32(&tail->full[s]);
        const uint32_t dst = ring_base + s * kStageBytes;
        if (elect_one()) {
          // only the pair leader arrives (once, expecting every byte that lands in its pair for this stage); bytes
          // may land first and drive the tx-count negative - the phase cannot complete before the arrive
          if (c == 0) mbar_arrive_expect_tx(full, 2 * kStageBytes);
          if (kShareC) {
            tma_load_2d_2cta(dst, &tmap_q128, full, ks * kBlockK, q_row0);
            tma_load_2d_2cta_mc(dst + kPanelBytes + pr * kBoxBytes, &tmap_c64, full, ks * kBlockK,
                                c_row0 + pr * (kHalfN / 2), mc_mask);
          } else {
            tma_load_2d_2cta_mc(dst + pr * kBoxBytes, &tmap_q64, full, ks * kBlockK, q_row0 + pr * (kBlockM / 2),
                                mc_mask);
            tma_load_2d_2cta(dst + kPanelBytes, &tmap_c128, full, ks * kBlockK, c_row0);
          }
        }
        __syncwarp();
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
    }
    if (rank == 0 && lane == 0 && n_cols > 1) prog[col] = 0x7FFFFFFF;   // done: never hold a peer back
  } else if (warp == 1) {
    // ===================== MMA issuer (each pair's leader CTA; whole warp, elected issue) =====================
    if (c == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, kBlockN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(smem_u32(&tail->tmem_empty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kBlockN);
        for (int ks = 0; ks < p.num_kb; ++ks) {
          mbar_wait(smem_u32(&tail->full[s]), ph);
          tc_fence_after();
          const uint32_t st = ring_base + s * kStageBytes;
          if (elect_one()) {
            const uint64_t adesc0 = make_sw128_kmajor_desc(st);
            const uint64_t bdesc0 = make_sw128_kmajor_desc(st + kPanelBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)   // +32 bytes per k-step = +2 in the 16-byte address field
              umma_bf16_ss_2cta(d_tmem, adesc0 + static_cast<uint64_t>(2 * k), bdesc0 + static_cast<uint64_t>(2 * k),
                                idesc, (ks | k) != 0 ? 1u : 0u);
            umma_commit_2cta_mask(smem_u32(&tail->empty[s]), 0xF);                 // frees the slot in all four CTAs
            if (ks == p.num_kb - 1) umma_commit_2cta_mask(smem_u32(&tail->tmem_full[as]), pair_mask);
          }
          __syncwarp();
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: thread <-> query (all four CTAs) =====================
    run_epilogue<true, 2>(p, tail->invc, tail->tmem_full, tail->tmem_empty, tmem_base, qb, unit, static_cast<uint32_t>(c),
                          t_first, t1, warp, lane, t_step, n_iter);
  }

  tc_fence_before();
  cluster_sync_all();   // every CTA's smem/TMEM must outlive the last remote MMA read, multicast write and arrive
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

constexpr size_t kMaxSmem = 232448;   // 227 KiB opt-in limit per CTA on sm_100

template <bool kShareC>
cudaError_t launch_t(const CUtensorMap& q128, const CUtensorMap& q64, const CUtensorMap& c128, const CUtensorMap& c64,
                     const ScanParams& p, int n_clusters, size_t smem, int n_stages, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(scan4_kernel<kShareC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  scan4_kernel<kShareC><<<4 * n_clusters, 64 + 128 * 2, smem, stream>>>(q128, q64, c128, c64, p, n_stages);
  return cudaGetLastError();
}

}  // namespace

// How many 4-CTA clusters of this kernel can be resident at once (one CTA per SM; GPCs whose SM count is not a
// multiple of 4 leave SMs unused).  Cached per device; 0 on error.
int scan4_max_clusters(bool smem_aligned) {
  static int cached[64][2];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int& slot = cached[dev][smem_aligned ? 1 : 0];
  if (slot == 0) {
    const int n_stages = smem_aligned ? kMaxStages4 : kMaxStages4 - 1;
    size_t smem = static_cast<size_t>(n_stages) * kStageBytes + sizeof(SmemTail4);
    if (smem + 1024 <= kMaxSmem) smem += 1024;
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(64 + 128 * 2);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaFuncSetAttribute(scan4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
            cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, scan4_kernel<true>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = -1;
    }
    slot = n > 0 ? n : -1;
  }
  return slot > 0 ? slot : 0;
}

// p.QB 256-query blocks; p.RC corpus ranges (clusters per column); p.R list units per query block.
// share_c requires an even p.QB.  Grid = 4 * n_clusters CTAs, n_clusters = p.RC * (share_c ? QB / 2 : QB).
cudaError_t launch_scan4(const CUtensorMap& tmap_q128, const CUtensorMap& tmap_q64, const CUtensorMap& tmap_c128,
                         const CUtensorMap& tmap_c64, const ScanParams& p, bool share_c, bool smem_aligned,
                         cudaStream_t stream, int* ring_stages_out) {
  const int n_stages = smem_aligned ? kMaxStages4 : kMaxStages4 - 1;
  size_t smem = static_cast<size_t>(n_stages) * kStageBytes + sizeof(SmemTail4);
  if (smem + 1024 <= kMaxSmem) smem += 1024;
  if (ring_stages_out) *ring_stages_out = n_stages;
  const int n_clusters = p.RC * (share_c ? p.QB / 2 : p.QB);
  return share_c ? launch_t<true>(tmap_q128, tmap_q64, tmap_c128, tmap_c64, p, n_clusters, smem, n_stages, stream)
                 : launch_t<false>(tmap_q128, tmap_q64, tmap_c128, tmap_c64, p, n_clusters, smem, n_stages, stream);
}

}  // namespace rbk

#endif  // RBK_EXPERIMENTAL
