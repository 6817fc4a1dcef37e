// Prints how many thread-block clusters of size 2 / 4 / 8 the device can hold at once for a 1-CTA-per-SM kernel
// (200 KB dynamic smem, 256 threads) - decides whether cluster-4 TMA multicast is worth building (GPC packing).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) dummy(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64, 1, 1);
        cfg.blockDim = dim3(256, 1, 1);
        cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster_size %d: max active clusters %d (= %d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("SMs %d\n", pr.multiProcessorCount);
    return 0;
}
