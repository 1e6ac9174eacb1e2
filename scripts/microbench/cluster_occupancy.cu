// How many thread-block clusters of 1 / 2 / 4 / 8 CTAs with the resources of gemm_i8_mod_kernel (320 threads, 193 KiB dynamic shared
// memory: one CTA per SM) can be co-resident on this GPU?  (cudaOccupancyMaxActiveClusters; GPC boundaries limit larger clusters.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    const int smem = 4 * 49152 + 1024 + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    for (int cs = 1; cs <= 16; cs *= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 148, 1, 1); cfg.blockDim = dim3(320, 1, 1); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
