// rbk_scan2.cu — K1 for batches of more than 128 queries: the same fused scan as
// rbk_scan.cu, on CTA PAIRS (thread-block cluster of 2, tcgen05 cta_group::2).
//
// Why: at B >= 256 the 1-CTA kernel is bound by L2->SM bandwidth (ncu, profiles/): every
// 128-query CTA pulls the whole 256-row corpus tile plus its query slab through L2 for each
// tile.  A pair computes a 256-query x 256-row tile with ONE copy of the corpus tile split
// across the two SMs (each CTA stages 128 corpus rows and its own 128 queries; the UMMA
// reads both halves), which halves the corpus bytes per query row.
//
// Protocol (leader = even CTA of the pair):
//   * both CTAs' TMA loads complete on the LEADER's full[s] barrier (count 1: the leader's
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

constexpr int kStages2 = 6;
constexpr int kHalfN = kBlockN / 2;                  // corpus rows staged per CTA
constexpr int kA2Bytes = kBlockM * kBlockK * 2;      // 16 KiB
constexpr int kB2Bytes = kHalfN * kBlockK * 2;       // 16 KiB
constexpr int kStage2Bytes = kA2Bytes + kB2Bytes;    // 32 KiB per CTA per stage
constexpr int kTmemCols = 512;
constexpr int kEpiThreads = 128;

struct SmemTail2 {
  float invc[2][kBlockN];
  unsigned long long full[kStages2];
  unsigned long long empty[kStages2];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kScanThreads, 1)
scan2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
             const ScanParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  SmemTail2* tail = reinterpret_cast<SmemTail2*>(smem + kStages2 * kStage2Bytes);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int qb = pair % p.QB;                        // 256-query block
  const int r = pair / p.QB;                         // corpus range
  const int t0 = static_cast<int>(static_cast<long long>(p.n_tiles) * r / p.R);
  const int t1 = static_cast<int>(static_cast<long long>(p.n_tiles) * (r + 1) / p.R);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);
      mbar_init(smem_u32(&tail->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tail->tmem_full[a]), 1);
      mbar_init(smem_u32(&tail->tmem_empty[a]), 8);
    }
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
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      volatile int* prog = p.progress + r * p.QB;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t0; tile < t1; ++tile) {
        const int it = tile - t0;
        if (rank == 0 && p.QB > 1 && (it & 1) == 0) {   // bounded-lag lockstep, see rbk_scan.cu
          prog[qb] = it;
          for (int o = 0; o < p.QB; ++o) {
            if (o == qb) continue;
            const long long w0 = clock64();
            while (prog[o] < it - kMaxLeadTiles) {
              __nanosleep(200);
              if (clock64() - w0 > (1ll << 24)) break;
            }
          }
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
          const uint32_t full = smem_u32(&tail->full[s]);
          // Only the leader arrives (once, expecting BOTH CTAs' bytes).  The peer's bytes may land
          // first and drive the tx-count negative; the phase cannot complete before the leader's
          // arrive, and the peer re-uses a slot only after the leader's MMAs consumed it.
          if (rank == 0) mbar_arrive_expect_tx(full, 2 * kStage2Bytes);
          const uint32_t a_dst = smem_base + s * kStage2Bytes;
          tma_load_2d_2cta(a_dst, &tmap_q, full, kb * kBlockK, qb * 2 * kBlockM + static_cast<int>(rank) * kBlockM);
          tma_load_2d_2cta(a_dst + kA2Bytes, &tmap_c, full, kb * kBlockK,
                           tile * kBlockN + static_cast<int>(rank) * kHalfN);
          if (++s == kStages2) { s = 0; ph ^= 1u; }
        }
      }
      if (rank == 0 && p.QB > 1) prog[qb] = 0x7FFFFFFF;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, kBlockN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = t0; tile < t1; ++tile) {
        mbar_wait(smem_u32(&tail->tmem_empty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kBlockN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&tail->full[s]), ph);
          tc_fence_after();
          const uint32_t a0 = smem_base + s * kStage2Bytes;
          const uint32_t b0 = a0 + kA2Bytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_bf16_ss_2cta(d_tmem, make_sw128_kmajor_desc(a0 + k * 32), make_sw128_kmajor_desc(b0 + k * 32),
                              idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(smem_u32(&tail->empty[s]));
          if (kb == p.num_kb - 1) umma_commit_2cta(smem_u32(&tail->tmem_full[as]));
          if (++s == kStages2) { s = 0; ph ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: thread <-> query (both CTAs) =====================
    const int quad = warp & 3;
    const int qrow = quad * 32 + lane;
    const int qin = static_cast<int>(rank) * kBlockM + qrow;   // row inside the 256-query block
    const int q = qb * 2 * kBlockM + qin;
    const bool q_valid = q < p.B;
    const int et = threadIdx.x - 64;
    FilterState fs;
    filter_init(fs, q_valid, q_valid ? p.thr_init[q] : INFINITY, q_valid ? p.inv_norm_q[q] : 0.f,
                p.cand + (static_cast<size_t>(qb * p.R + r) * (2 * kBlockM) + qin) * static_cast<size_t>(kListCap),
                p.hist + static_cast<size_t>(q_valid ? q : 0) * kHistBins, p.maxbin + (q_valid ? q : 0));
    int as = 0;
    uint32_t aph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      const int row0 = tile * kBlockN;
      const int it = tile - t0;
      tail->invc[as][et] = __ldg(p.inv_norm_c + row0 + et);
      tail->invc[as][kEpiThreads + et] = __ldg(p.inv_norm_c + row0 + kEpiThreads + et);
      named_bar_sync(1, kEpiThreads);
      if (it != 0 && (it < 8 || (it & 3) == 0)) filter_refresh(fs, p.kprime);   // overlaps this tile's MMAs
      mbar_wait(smem_u32(&tail->tmem_full[as]), aph);
      tc_fence_after();
      const float* invc = tail->invc[as];
#pragma unroll 1
      for (int chunk = 0; chunk < kBlockN / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(as * kBlockN + chunk * 32),
                           v);
        tmem_wait_ld();
        filter_chunk(fs, v, invc + chunk * 32, static_cast<uint32_t>(row0 + chunk * 32));
        if (p.dbg_scores != nullptr && q_valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int row = row0 + chunk * 32 + j;
            if (row < p.n_rows)
              p.dbg_scores[static_cast<size_t>(q) * p.n_rows + row] = __uint_as_float(v[j]) * invc[chunk * 32 + j];
          }
        }
        filter_compact_if_needed(fs, p.kprime, lane);
        if (it == 0) filter_refresh(fs, p.kprime);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&tail->tmem_empty[as]));
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
    p.cand_cnt[(qb * p.R + r) * (2 * kBlockM) + qin] = fs.cnt;
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's smem/TMEM must outlive the leader's last MMA and all remote arrives
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

}  // namespace

size_t scan2_smem_bytes() { return static_cast<size_t>(kStages2) * kStage2Bytes + sizeof(SmemTail2) + 1024; }

// p.QB counts 256-query blocks, p.R CTA pairs per block; grid = 2 * QB * R CTAs.
cudaError_t launch_scan2(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c_half, const ScanParams& p,
                         cudaStream_t stream) {
  const size_t smem = scan2_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(scan2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan2_kernel<<<2 * p.QB * p.R, kScanThreads, smem, stream>>>(tmap_q, tmap_c_half, p);
  return cudaGetLastError();
}

}  // namespace rbk
