// Developer probe: how many clusters of 2 / 4 / 8 CTAs with ~227 KB of shared memory can be resident at once?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 231936);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 231936;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %d (= %d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
