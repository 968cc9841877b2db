// tma_stream_probe.cu — how fast can 148 persistent CTAs stream a bf16 corpus from HBM through TMA into shared memory,
// (a) as the scan kernels do today: row-major [N][768], 12 boxes of 256 rows x 64 columns per tile (256 strided
//     128-byte pieces per box), vs (b) tile-major [N/256][12][256][64]: every box one contiguous 32 KB read?
// Decides whether re-laying the corpus out tile-major is worth it for the HBM-bound (B <= 256) shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_probe tma_stream_probe.cu && ./tma_stream_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int kStages = 6, kStageBytes = 32768, kKB = 12;

// mode 0: row-major, box at (col kb*64, row tile*256).  mode 1: tile-major, box at (0, (tile*12+kb)*256).
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmap, int n_tiles, int mode,
                                                       unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[kStages], empty[kStages];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int t0 = static_cast<int>(static_cast<long long>(n_tiles) * blockIdx.x / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(n_tiles) * (blockIdx.x + 1) / gridDim.x);
  if (threadIdx.x == 0) {          // producer
    int s = 0; uint32_t ph = 0;
    for (int tile = t0; tile < t1; ++tile)
      for (int kb = 0; kb < kKB; ++kb) {
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        mbar_expect(smem_u32(&full[s]), kStageBytes);
        if (mode == 0) tma_load_2d(base + s * kStageBytes, &tmap, smem_u32(&full[s]), kb * 64, tile * 256);
        else tma_load_2d(base + s * kStageBytes, &tmap, smem_u32(&full[s]), 0, (tile * kKB + kb) * 256);
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
  } else if (threadIdx.x == 32) {  // consumer: free the slot as soon as it is full
    int s = 0; uint32_t ph = 0; unsigned long long acc = 0;
    for (int tile = t0; tile < t1; ++tile)
      for (int kb = 0; kb < kKB; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        acc += *reinterpret_cast<volatile unsigned long long*>(smem + (base - smem_u32(smem)) + s * kStageBytes);
        mbar_arrive(smem_u32(&empty[s]));
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    if (acc == 0x1234567ull) *sink = acc;
  }
}

int main() {
  const long long n_rows = 4000000 / 256 * 256;   // 6.1 GB
  const int d = 768, n_tiles = static_cast<int>(n_rows / 256);
  void* buf; cudaMalloc(&buf, static_cast<size_t>(n_rows) * d * 2); cudaMemset(buf, 1, static_cast<size_t>(n_rows) * d * 2);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  CUtensorMap maps[2];
  {
    cuuint64_t gdim[2] = {768, (cuuint64_t)n_rows}, gstr[1] = {768 * 2}; cuuint32_t box[2] = {64, 256}, es[2] = {1, 1};
    CUresult r = enc(&maps[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode 0 failed %d\n", (int)r);
  }
  {
    cuuint64_t gdim[2] = {64, (cuuint64_t)n_rows * 12}, gstr[1] = {128}; cuuint32_t box[2] = {64, 256}, es[2] = {1, 1};
    CUresult r = enc(&maps[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode 1 failed %d\n", (int)r);
  }
  const int smem = kStages * kStageBytes + 1024;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const double bytes = static_cast<double>(n_rows) * d * 2;
  for (int rep = 0; rep < 3; ++rep)
    for (int mode = 0; mode < 2; ++mode) {
      cudaEventRecord(a);
      stream_kernel<<<148, 64, smem>>>(maps[mode], n_tiles, mode, sink);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("%s: %.3f ms  %.0f GB/s  (%s)\n", mode ? "tile-major contiguous 32 KB boxes" : "row-major 256 x 128 B strided boxes ", ms,
             bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
