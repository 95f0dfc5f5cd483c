// Prints how many clusters of size 1,2,4,8 of a 1-CTA-per-SM kernel (given dynamic shared memory) can be co-resident.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("SMs %d\n", prop.multiProcessorCount);
    for (int smem : {200 * 1024, 224 * 1024}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int cs : {1, 2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(608); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("smem %d KB cluster %2d: max active clusters %d (%d SMs) %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
