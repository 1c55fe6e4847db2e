// How many thread-block clusters of a given size can a B200 keep resident with ~1 CTA per SM (big dynamic shared memory)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy tools/cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t c; memset(&c, 0, sizeof(c));
        c.gridDim = dim3(cs * 64); c.blockDim = dim3(288); c.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        c.attrs = a; c.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &c);
        printf("cluster size %2d: %3d resident clusters = %3d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
