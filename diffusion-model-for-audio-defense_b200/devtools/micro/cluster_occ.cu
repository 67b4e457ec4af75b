// Development aid: how many thread-block clusters of 1 / 2 / 4 / 8 CTAs (one CTA per SM: 224 KB of shared memory, like k1_layer)
// can be resident on this GPU at once?  148 SMs in GPCs of unequal size: a cluster must fit inside one GPC.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occ cluster_occ.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) dummy(int* p) {
  extern __shared__ int sm[];
  if (p) p[0] = sm[0];
}
int main() {
  const int smem = 224 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  printf("%s: %d SMs\n", pr.name, pr.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 148), cfg.blockDim = dim3(384), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d CTAs resident of %d SMs (%s)\n", cs, n, n * cs, pr.multiProcessorCount,
           cudaGetErrorString(e));
  }
  return 0;
}
