// rbk_scan3.cu — K1, CTA pairs with the QUERY OPERAND IN TENSOR MEMORY (dim <= 768).
//
// Why: the streamed pair kernel (rbk_scan2.cu) is bound by L2->SM bandwidth (~9-10 TB/s,
// profiles/r01_scan_cfg3_v4.txt) and half of those bytes are the query slab, re-read for every
// corpus tile.  Keeping the queries in shared memory removes them from the stream but leaves
// only 32 KB for the corpus ring (latency-bound, measured slower).  Tensor memory has room:
// a CTA's 128 queries x 768 bf16 are 128 lanes x 384 columns, and tcgen05.mma accepts its A
// operand from TMEM.  The remaining 128 columns hold two 64-column accumulators, and ALL of
// shared memory (24 stages x 8 KB per CTA) is a deep ring for the only thing that still
// streams: the corpus.  L2->SM bytes per 256 queries x 256 rows drop from 786 KB to 393 KB.
//
//   TMEM columns   [0, dpad/2)      queries: lane = query row, column c = elements 2c, 2c+1
//                  [384, 448) [448, 512)   accumulators (double buffered), N = 64
//   MMA            tcgen05.mma.cta_group::2.kind::f16  M = 256 (128 per CTA), N = 64, K = 16
//                  A from TMEM, B = 32 corpus rows per CTA from smem (SWIZZLE_128B)
//   sub-tile       64 corpus rows; a 256-row tile = 4 sub-tiles (lockstep / ranges as before)
//
// The pair protocol (leader-only arrive.expect_tx, multicast commits, remote tmem_empty
// arrive) is the one of rbk_scan2.cu.
#ifdef RBK_EXPERIMENTAL   // measured and rejected: not part of the default build
#include "rbk_epilogue.cuh"
#include "rbk_internal.h"
#include "rbk_ptx.cuh"

namespace rbk {

namespace {

#ifndef RBK_TS_SUBN
#define RBK_TS_SUBN 64
#endif
constexpr int kSubN = RBK_TS_SUBN;              // corpus rows per accumulator (UMMA N): 64 (2 buffers) or 128 (1)
constexpr int kAccStages = kSubN == 64 ? 2 : 1;
constexpr int kSubHalf = kSubN / 2;             // rows staged per CTA
#ifndef RBK_TS_KB_PER_STAGE
#define RBK_TS_KB_PER_STAGE 4
#endif
constexpr int kKbPerStage = RBK_TS_KB_PER_STAGE; // 64-column k-blocks per ring stage (16 MMAs = 512 tensor cycles at 4)
constexpr int kKbBytes = kSubHalf * kBlockK * 2;            // 4 KiB: 32 rows x 64 bf16
constexpr int kStage3Bytes = kKbPerStage * kKbBytes;        // 8 KiB per CTA per stage
constexpr int kStages3 = 192 * 1024 / kStage3Bytes;         // 192 KiB ring per CTA
constexpr int kTmemCols = 512;
constexpr int kAccCol0 = 384;                               // accumulators live above the queries
constexpr int kEpi = 128;

struct SmemTail3 {
  float invc[2][kBlockN];   // per 256-row tile, double buffered by tile parity
  unsigned long long full[kStages3];
  unsigned long long empty[kStages3];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  unsigned long long a_ready;
  uint32_t tmem_base;
};

// D[tmem, both CTAs] (+)= A[tmem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T
__device__ __forceinline__ void umma_bf16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective: lane t writes 32 consecutive 32-bit columns of TMEM lane (taddr.lane + t).
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]),
        "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kScanThreads, 1)
scan3_kernel(const __grid_constant__ CUtensorMap tmap_c, const ScanParams p, const uint16_t* __restrict__ q_bf16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  SmemTail3* tail = reinterpret_cast<SmemTail3*>(smem + kStages3 * kStage3Bytes);
  const uint32_t ring_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int qb = pair % p.QB;
  const int r = pair / p.QB;
  const int t0 = static_cast<int>(static_cast<long long>(p.n_tiles) * r / p.R_local);
  const int t1 = static_cast<int>(static_cast<long long>(p.n_tiles) * (r + 1) / p.R_local);
  const int n_stages_per_sub = p.num_kb / kKbPerStage;   // host guarantees num_kb % kKbPerStage == 0
  constexpr int kSubPerTile = kBlockN / kSubN;           // 4

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < kStages3; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);
      mbar_init(smem_u32(&tail->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tail->tmem_full[a]), 1);
      mbar_init(smem_u32(&tail->tmem_empty[a]), 8);
    }
    mbar_init(smem_u32(&tail->a_ready), 8);
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
    // ===================== TMA producer: corpus only (whole warp, elected issue) =====================
    volatile int* prog = p.progress + r * p.QB;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      if (rank == 0 && lane == 0) lockstep_pace(prog, p.QB, qb, tile - t0, p.max_lead_tiles);
      __syncwarp();
      for (int j = 0; j < kSubPerTile; ++j) {
        const int c_row0 = tile * kBlockN + j * kSubN + static_cast<int>(rank) * kSubHalf;
        for (int st = 0; st < n_stages_per_sub; ++st) {
          mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
          const uint32_t full = smem_u32(&tail->full[s]);
          const uint32_t dst = ring_base + s * kStage3Bytes;
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(full, 2 * kStage3Bytes);
#pragma unroll
            for (int kk = 0; kk < kKbPerStage; ++kk)
              tma_load_2d_2cta(dst + kk * kKbBytes, &tmap_c, full, (st * kKbPerStage + kk) * kBlockK, c_row0);
          }
          __syncwarp();
          if (++s == kStages3) { s = 0; ph ^= 1u; }
        }
      }
    }
    if (rank == 0 && lane == 0 && p.QB > 1) prog[qb] = 0x7FFFFFFF;
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp, elected issue) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, kSubN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      mbar_wait(smem_u32(&tail->a_ready), 0u);   // both CTAs' queries are in TMEM
      tc_fence_after();
      for (int tile = t0; tile < t1; ++tile) {
        for (int j = 0; j < kSubPerTile; ++j) {
          mbar_wait(smem_u32(&tail->tmem_empty[as]), aph ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(kAccCol0 + as * kSubN);
          for (int st = 0; st < n_stages_per_sub; ++st) {
            mbar_wait(smem_u32(&tail->full[s]), ph);
            tc_fence_after();
            const uint64_t bdesc0 = make_sw128_kmajor_desc(ring_base + s * kStage3Bytes);
            const uint32_t a_col0 = tmem_base + static_cast<uint32_t>(st * kKbPerStage * (kBlockK / 16) * 8);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < kKbPerStage; ++kk) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  const int kl = kk * (kBlockK / 16) + k;   // k-step inside the stage: 16 elements = 8 TMEM columns
                  // descriptor start-address field counts 16-byte units: + (kk * 4 KiB + k * 32 B) / 16
                  umma_bf16_ts_2cta(d_tmem, a_col0 + static_cast<uint32_t>(kl * 8),
                                    bdesc0 + static_cast<uint64_t>((kk * kKbBytes + k * 32) >> 4), idesc,
                                    (st | kl) != 0 ? 1u : 0u);
                }
              }
              umma_commit_2cta(smem_u32(&tail->empty[s]));
              if (st == n_stages_per_sub - 1) umma_commit_2cta(smem_u32(&tail->tmem_full[as]));
            }
            __syncwarp();
            if (++s == kStages3) { s = 0; ph ^= 1u; }
          }
          if (++as == kAccStages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue warps: load queries into TMEM, then filter =====================
    const int quad = warp & 3;
    const int qrow = quad * 32 + lane;
    const int qin = static_cast<int>(rank) * kBlockM + qrow;
    const int q = qb * 2 * kBlockM + qin;
    const bool q_valid = q < p.B;
    const int et = threadIdx.x - 64;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    {
      // query row -> TMEM lane: 32 columns (64 bf16, 128 bytes) per tcgen05.st
      const uint4* src = reinterpret_cast<const uint4*>(q_bf16 + static_cast<size_t>(q_valid ? q : 0) * p.dpad);
      const int n_groups = (p.dpad + 63) / 64;          // groups of 64 elements
      const int n_u4 = p.dpad / 8;                       // dpad is a multiple of 8
      for (int g = 0; g < n_groups; ++g) {
        uint32_t v[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint4 w = make_uint4(0u, 0u, 0u, 0u);
          if (q_valid && g * 8 + u < n_u4) w = __ldg(src + g * 8 + u);
          v[4 * u + 0] = w.x;
          v[4 * u + 1] = w.y;
          v[4 * u + 2] = w.z;
          v[4 * u + 3] = w.w;
        }
        tmem_st_32x32b_x32(lane_base + static_cast<uint32_t>(g * 32), v);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&tail->a_ready));
    }
    FilterState fs;
    filter_init(fs, q_valid, q_valid ? p.thr_init[q] : INFINITY, q_valid ? p.inv_norm_q[q] : 0.f,
                p.cand + (static_cast<size_t>(qb * p.R + r) * (2 * kBlockM) + qin) * static_cast<size_t>(kListCap),
                p.hist + static_cast<size_t>(q_valid ? q : 0) * kHistBins, p.maxbin + (q_valid ? q : 0));
    int as = 0;
    uint32_t aph = 0;
    // 1/||c|| of the NEXT tile travels in registers while the current tile is processed, so its
    // L2/HBM latency never sits between two 64-row sub-tiles (it did: ~1 us per sub-tile).
    float nx0 = 0.f, nx1 = 0.f;
    if (t0 < t1) {
      nx0 = __ldg(p.inv_norm_c + t0 * kBlockN + et);
      nx1 = __ldg(p.inv_norm_c + t0 * kBlockN + kEpi + et);
    }
    for (int tile = t0; tile < t1; ++tile) {
      const int it = tile - t0;
      float* invc_tile = tail->invc[it & 1];
      invc_tile[et] = nx0;
      invc_tile[kEpi + et] = nx1;
      if (tile + 1 < t1) {
        nx0 = __ldg(p.inv_norm_c + (tile + 1) * kBlockN + et);
        nx1 = __ldg(p.inv_norm_c + (tile + 1) * kBlockN + kEpi + et);
      }
      named_bar_sync(1, kEpi);
      if (refresh_due(it)) filter_refresh(fs, p.kprime);
      for (int j = 0; j < kSubPerTile; ++j) {
        const int row0 = tile * kBlockN + j * kSubN;
        const float* invc = invc_tile + j * kSubN;
        mbar_wait(smem_u32(&tail->tmem_full[as]), aph);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = 0; chunk < kSubN / 32; ++chunk) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_base + static_cast<uint32_t>(kAccCol0 + as * kSubN + chunk * 32), v);
          tmem_wait_ld();
          filter_chunk(fs, v, invc + chunk * 32, static_cast<uint32_t>(row0 + chunk * 32));
          if (p.dbg_scores != nullptr && q_valid) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int row = row0 + chunk * 32 + c;
              if (row < p.n_rows)
                p.dbg_scores[static_cast<size_t>(q) * p.n_rows + row] = __uint_as_float(v[c]) * invc[chunk * 32 + c];
            }
          }
          filter_compact_if_needed(fs, p.kprime, lane);
          if (it == 0 && t1 - t0 > 2) filter_refresh(fs, p.kprime);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(smem_u32(&tail->tmem_empty[as]));
        if (++as == kAccStages) { as = 0; aph ^= 1u; }
      }
    }
    p.cand_cnt[(qb * p.R + r) * (2 * kBlockM) + qin] = fs.cnt;
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

}  // namespace

int scan3_box_rows() { return kSubHalf; }

// Queries fit TMEM next to two 64-column accumulators, and the ring stages hold whole k-block pairs.
bool scan3_fits(int dpad) {
  const int num_kb = (dpad + kBlockK - 1) / kBlockK;
  return dpad <= 2 * kAccCol0 && num_kb % kKbPerStage == 0;
}

// tmap_c: 32-row x 64-column boxes, SWIZZLE_128B.  p.QB counts 256-query blocks, p.R pairs per block.
cudaError_t launch_scan3(const CUtensorMap& tmap_c, const ScanParams& p, const uint16_t* q_bf16,
                         cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(kStages3) * kStage3Bytes + sizeof(SmemTail3) + 1024;
  cudaError_t e = cudaFuncSetAttribute(scan3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan3_kernel<<<2 * p.QB * p.R, kScanThreads, smem, stream>>>(tmap_c, p, q_bf16);
  return cudaGetLastError();
}

}  // namespace rbk

#endif  // RBK_EXPERIMENTAL
