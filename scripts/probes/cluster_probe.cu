// cluster_probe.cu — how many thread-block clusters of 2 / 4 / 8 CTAs (one CTA per SM, ~227 KB smem each) can be
// co-resident on this GPU?  Decides the grid of the cluster-multicast scan kernel.  Also prints which SMs a
// cluster-of-4 grid actually lands on.   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void probe_kernel(int* smid) {
  extern __shared__ char smem[];
  if (threadIdx.x == 0) {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    smid[blockIdx.x] = static_cast<int>(s);
    smem[0] = 1;
  }
  // stay resident long enough that every cluster of the grid must be co-scheduled or wait
  const long long t0 = clock64();
  while (clock64() - t0 < 2000000) {
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smem = 227 * 1024 - 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)  [%s]\n", cs, n, n * cs, cudaGetErrorString(e));
    if (e != cudaSuccess) { cudaGetLastError(); continue; }
    if (cs == 4 || cs == 2) {
      int* d;
      cudaMalloc(&d, sizeof(int) * n * cs);
      cfg.gridDim = dim3(n * cs);
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      e = cudaLaunchKernelEx(&cfg, probe_kernel, d);
      cudaEventRecord(b);
      cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      std::vector<int> h(n * cs);
      cudaMemcpy(h.data(), d, sizeof(int) * n * cs, cudaMemcpyDeviceToHost);
      std::vector<int> used(prop.multiProcessorCount, 0);
      for (int v : h) if (v >= 0 && v < prop.multiProcessorCount) used[v]++;
      int distinct = 0, multi = 0;
      for (int v : used) { distinct += v > 0; multi += v > 1; }
      printf("  launched %d CTAs: %d distinct SMs, %d SMs ran more than one CTA, %.3f ms (one wave = ~1.1 ms) [%s]\n",
             n * cs, distinct, multi, ms, cudaGetErrorString(e));
      printf("  idle SMs:");
      for (int i = 0; i < prop.multiProcessorCount; ++i) if (!used[i]) printf(" %d", i);
      printf("\n");
      cudaFree(d);
    }
  }
  return 0;
}
