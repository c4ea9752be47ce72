// How many thread-block clusters of each size are resident at once on this GPU for a one-CTA-per-SM kernel (210 KB of dynamic
// shared memory, 384 threads) -- what a multicast GEMM mode can occupy.   nvcc -arch=sm_100a -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cl : {1, 2, 3, 4, 6, 8, 12, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = 210 * 1024; cfg.gridDim = dim3(prop.multiProcessorCount / cl * cl);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: %3d clusters = %3d SMs  (%s)\n", cl, n, n * cl, cudaGetErrorString(e));
    cudaGetLastError();
  }
  return 0;
}
