// How many thread-block clusters of a given size can be co-resident for a 1-CTA-per-SM kernel (416 threads, ~226 KB of
// dynamic shared memory -- the footprint of the decoder's GEMM kernels)?  Build: nvcc -arch=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(416, 1) probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = (int)smem[0];
}
int main() {
  const int smem = 226 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
  const int sizes[] = {2, 4, 6, 8, 10, 12, 14, 16};
  for (int cs : sizes) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 16); cfg.blockDim = dim3(416); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)%s%s\n", cs, n, n * cs, e == cudaSuccess ? "" : "  error: ", e == cudaSuccess ? "" : cudaGetErrorString(e));
    // and an actual launch
    e = cudaLaunchKernelEx(&cfg, probe_kernel, (int*)nullptr);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) printf("   launch: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaGetLastError();
  }
  return 0;
}
