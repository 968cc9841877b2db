// rbk_scan.cu — K1: fused  Q x C^T (tcgen05, bf16 -> fp32 in TMEM)  +  row-norm scaling
//               +  per-query running top-k' selection in the epilogue.
//
// Replaces the hot loop of VectorStore.search (reference src/knowledge/store/
// vector-store.ts:207-215 calling cosineSimilarity, src/knowledge/indexer/
// embedder.ts:168-184) for a whole batch of queries at once.  The score matrix never
// reaches HBM: each epilogue thread owns one query (one TMEM lane), streams the 256
// scores of a tile out of TMEM, compares them with that query's running threshold and
// appends the rare survivors to a small per-(CTA, query) candidate list.
//
// Work decomposition (persistent CTAs, one per SM):
//   CTA c -> query block qb = c % QB (128 queries), corpus range r = c / QB.
//   CTAs that share r walk the same corpus tiles at the same pace, so a tile is pulled
//   from HBM once and served to the other query blocks from L2.
// Warp roles: warp 0 = TMA producer (1 thread), warp 1 = TMEM owner + MMA issuer
//   (1 thread), warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM double buffer full/empty
//   (MMA <-> epilogue): the epilogue of tile i overlaps the MMAs of tile i+1.
#include "rbk_epilogue.cuh"
#include "rbk_internal.h"
#include "rbk_ptx.cuh"

namespace rbk {

namespace {

constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KiB  query k-slab
constexpr int kBBytes = kBlockN * kBlockK * 2;      // 32 KiB  corpus k-slab
constexpr int kStageBytes = kABytes + kBBytes;      // 48 KiB
constexpr int kTmemCols = 512;                      // 2 accumulator stages x 256 fp32 columns

struct SmemTail {
  float invc[2][kBlockN];  // 1/||c|| of the tile's rows, per accumulator stage
  unsigned long long full[kStages];
  unsigned long long empty[kStages];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
            const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024-B alignment
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + kStages * kStageBytes);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x % p.QB;
  const int r = blockIdx.x / p.QB;
  const int t0 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * r / p.R_local);
  const int t1 = p.tile_begin + static_cast<int>(static_cast<long long>(p.tile_count) * (r + 1) / p.R_local);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&tail->full[s]), 1);
      mbar_init(smem_u32(&tail->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tail->tmem_full[a]), 1);
      mbar_init(smem_u32(&tail->tmem_empty[a]), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tail->tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, elected issue) =====================
    volatile int* prog = p.progress + p.prog_base + r * p.QB;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      if (lane == 0) lockstep_pace(prog, p.QB, qb, tile - t0, p.max_lead_tiles);
      __syncwarp();
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&tail->full[s]);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(full, kStageBytes);
          tma_load_2d(a_dst, &tmap_q, full, kb * kBlockK, qb * kBlockM);
          tma_load_2d(a_dst + kABytes, &tmap_c, full, kb * kBlockK, tile * kBlockN);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    if (lane == 0 && p.QB > 1) prog[qb] = 0x7FFFFFFF;  // done: never hold a peer back
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, elected issue) =====================
    constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM, kBlockN);
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(smem_u32(&tail->tmem_empty[as]), aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kBlockN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(smem_u32(&tail->full[s]), ph);
        tc_fence_after();
        const uint64_t adesc0 = make_sw128_kmajor_desc(smem_base + s * kStageBytes);
        const uint64_t bdesc0 = make_sw128_kmajor_desc(smem_base + s * kStageBytes + kABytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)   // +32 bytes per k-step = +2 in the 16-byte address field
            umma_bf16_ss(d_tmem, adesc0 + static_cast<uint64_t>(2 * k), bdesc0 + static_cast<uint64_t>(2 * k), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&tail->empty[s]));  // smem slot reusable once these MMAs retire
          if (kb == p.num_kb - 1) umma_commit(smem_u32(&tail->tmem_full[as]));
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  } else {
    // ===================== epilogue: thread <-> query =====================
    run_epilogue<false>(p, tail->invc, tail->tmem_full, tail->tmem_empty, tmem_base, qb, p.unit_base + r, 0u, t0, t1, warp,
                        lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

size_t scan_smem_bytes() { return static_cast<size_t>(kStages) * kStageBytes + sizeof(SmemTail) + 1024; }

cudaError_t launch_scan(const CUtensorMap& tmap_q, const CUtensorMap& tmap_c, const ScanParams& p,
                        cudaStream_t stream) {
  const size_t smem = scan_smem_bytes();
  // per-device attribute; cheap enough to set on every launch
  cudaError_t e = cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan_kernel<<<p.QB * p.R_local, kScanThreads, smem, stream>>>(tmap_q, tmap_c, p);
  return cudaGetLastError();
}

}  // namespace rbk
