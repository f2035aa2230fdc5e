// How many thread-block clusters of a k_track-shaped kernel (256 threads, ~124 KB dynamic shared memory, 1 CTA per SM) can be
// co-resident on this GPU, per cluster size?  Developer probe for the "rows pre-reduced in clusters" idea.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) k_probe(float * p)
{
    extern __shared__ float s[];
    s[threadIdx.x] = p ? p[threadIdx.x] : 0.f;
    __syncthreads();
    if(p) p[threadIdx.x] = s[255 - threadIdx.x];
}
int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    const size_t smem = 124 * 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for(int cs : {1, 2, 4, 8, 16})
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_probe, &cfg);
        printf("cluster size %2d: max active clusters %3d -> %3d CTAs co-resident (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
        // can a cooperative launch carry a cluster dimension?
        cfg.numAttrs = 2;
        cfg.gridDim = dim3((prop.multiProcessorCount / cs) * cs);
        float * p = nullptr;
        e = cudaLaunchKernelEx(&cfg, k_probe, p);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("   cooperative + cluster launch of %d CTAs: %s / %s\n", cfg.gridDim.x, cudaGetErrorString(e), cudaGetErrorString(e2));
        cudaGetLastError();
    }
    return 0;
}
