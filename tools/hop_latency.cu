// hop_latency.cu -- microbenchmark: latency of one flagged-chunk "hop" between two CTAs on different SMs
// (the synchronisation primitive of ef_track_kernel.cu).  CTA 0 sends epoch n, CTA k echoes it; the round trip
// measured with CTA 0's clock64 is two hops.  Variants: relaxed.gpu v4, volatile, 32-bit flag.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 * p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint4 * p, const uint4 & v)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed32(const unsigned * p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed32(unsigned * p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// mode 0: v4 relaxed; 1: u32 relaxed; 2: u32 volatile
__global__ void pingpong(uint4 * a, uint4 * b, int partner, int iters, int mode, long long * out)
{
    if(threadIdx.x != 0) return;
    if(blockIdx.x == 0)
    {
        long long t0 = clock64();
        for(int n = 1; n <= iters; n++)
        {
            if(mode == 0)
            {
                st_relaxed_v4(a, make_uint4(1, 2, 3, n));
                while(ld_relaxed_v4(b).w != (unsigned)n) {}
            }
            else if(mode == 1)
            {
                st_relaxed32((unsigned *)a, n);
                while(ld_relaxed32((unsigned *)b) != (unsigned)n) {}
            }
            else
            {
                *(volatile unsigned *)a = n;
                while(*(volatile unsigned *)b != (unsigned)n) {}
            }
        }
        out[0] = clock64() - t0;
    }
    else if((int)blockIdx.x == partner)
    {
        for(int n = 1; n <= iters; n++)
        {
            if(mode == 0)
            {
                while(ld_relaxed_v4(a).w != (unsigned)n) {}
                st_relaxed_v4(b, make_uint4(1, 2, 3, n));
            }
            else if(mode == 1)
            {
                while(ld_relaxed32((unsigned *)a) != (unsigned)n) {}
                st_relaxed32((unsigned *)b, n);
            }
            else
            {
                while(*(volatile unsigned *)a != (unsigned)n) {}
                *(volatile unsigned *)b = n;
            }
        }
    }
}

// fan-out: CTA 0 writes one flagged chunk per worker (W workers), every worker echoes into its own slot, CTA 0
// waits for all echoes with one thread per worker.  Round trip of the whole grid = what one tracker iteration pays.
__global__ void fan(uint4 * box, uint4 * echo, int iters, long long * out)
{
    const int W = gridDim.x - 1;
    if(blockIdx.x == 0)
    {
        long long t0 = clock64();
        for(int n = 1; n <= iters; n++)
        {
            for(int w = threadIdx.x; w < W; w += blockDim.x) st_relaxed_v4(box + w * 16, make_uint4(1, 2, 3, n));
            for(int w = threadIdx.x; w < W; w += blockDim.x)
                while(ld_relaxed_v4(echo + w).w != (unsigned)n) {}
            __syncthreads();
        }
        if(threadIdx.x == 0) out[0] = clock64() - t0;
    }
    else
    {
        if(threadIdx.x != 0) return;
        const int w = blockIdx.x - 1;
        for(int n = 1; n <= iters; n++)
        {
            while(ld_relaxed_v4(box + w * 16).w != (unsigned)n) {}
            st_relaxed_v4(echo + w, make_uint4(1, 2, 3, n));
        }
    }
}

int main()
{
    uint4 * buf;
    long long * out;
    cudaMalloc(&buf, 1 << 20);
    cudaMemset(buf, 0, 1 << 20);
    cudaMallocManaged(&out, 64);
    const int iters = 2000;
    for(int mode = 0; mode < 3; mode++)
        for(int partner : {1, 2, 37, 74, 100, 147})
        {
            cudaMemset(buf, 0, 1 << 20);
            void * args[] = {(void *)&buf, nullptr, (void *)&partner, (void *)&iters, (void *)&mode, (void *)&out};
            uint4 * b = buf + 64;
            args[1] = (void *)&b;
            cudaLaunchCooperativeKernel((const void *)pingpong, dim3(148), dim3(32), args, 0, 0);
            cudaDeviceSynchronize();
            printf("pingpong mode %d partner %3d: %.0f cycles per round trip (2 hops)\n", mode, partner, (double)out[0] / iters);
        }
    for(int threads : {160, 512})
    {
        cudaMemset(buf, 0, 1 << 20);
        uint4 * echo = buf + 16 * 256;
        void * args[] = {(void *)&buf, (void *)&echo, (void *)&iters, (void *)&out};
        cudaLaunchCooperativeKernel((const void *)fan, dim3(148), dim3(threads), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        printf("fan-out/in 147 workers, %d threads: %.0f cycles per round trip (%s)\n", threads, (double)out[0] / iters, cudaGetErrorString(e));
    }
    return 0;
}
