// allreduce_probe.cu -- microbenchmark for an ALTERNATIVE to the tracker kernel's gather -> solve -> broadcast chain: an
// order-independent all-reduce through L2 atomics.  Every CTA adds its partial sums as 64-bit fixed-point integers
// (value << 8 | 1: the low byte counts arrivals) into one accumulator line per quantity, announces itself on a replicated
// sentinel, polls its own sentinel replica, then reads the totals itself -- so every CTA could run the solve redundantly
// and no parameter hop back would be needed.  Measured: cycles per all-reduce round over all SMs, for several layouts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/allreduce_probe tools/allreduce_probe.cu && tools/allreduce_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kQ = 116;      // 58 quantities x (lo, hi)
constexpr int kRep = 32;     // sentinel replicas
constexpr int kIters = 48;   // rounds per launch (one pre-zeroed slot each)

__device__ __forceinline__ void red_add(unsigned long long * p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// stride_q: distance in 8-byte words between two quantities' accumulators (16 = one 128-byte line each, 1 = packed)
__global__ void __launch_bounds__(256, 1) probe(unsigned long long * acc, unsigned long long * sent, int stride_q, long long * out, int work_cycles)
{
    __shared__ unsigned long long s_tot[kQ];
    __shared__ int s_bad;
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    long long total = 0;
    unsigned long long check = 0;
    for(int it = 0; it < kIters; it++)
    {
        // some "pixel work" whose length differs a little from CTA to CTA
        const long long w0 = clock64();
        while(clock64() - w0 < work_cycles + (b % 7) * 20) {}
        __syncthreads();
        const long long t0 = clock64();
        unsigned long long * A = acc + (size_t)it * kQ * stride_q;
        unsigned long long * S = sent + (size_t)it * kRep * 16;
        if(t < kQ) red_add(A + (size_t)t * stride_q, ((unsigned long long)(b + t + 1) << 8) | 1ull);
        __syncthreads(); // every data atomic of this CTA has been issued (not necessarily performed) before its sentinels
        if(t >= 128 && t < 128 + kRep) red_add(S + (size_t)(t - 128) * 16, 1ull);
        if(t == 0)
        {
            const unsigned long long * mine = S + (size_t)(b % kRep) * 16;
            while(ld_relaxed(mine) != (unsigned long long)G) {}
        }
        __syncthreads();
        if(t == 0) s_bad = 0;
        __syncthreads();
        if(t < kQ)
        {
            unsigned long long v;
            int tries = 0;
            do
            {
                v = ld_relaxed(A + (size_t)t * stride_q);
                ++tries;
            } while((v & 0xffull) != (unsigned long long)G);
            if(tries > 1) atomicAdd(&s_bad, 1);
            s_tot[t] = v >> 8;
        }
        __syncthreads();
        const long long t1 = clock64();
        total += t1 - t0;
        check += s_tot[it % kQ] + s_bad * 0;
        if(t == 0 && b == 0 && s_bad) out[2] += s_bad; // totals that were not complete when the sentinel said so
    }
    if(t == 0)
    {
        if(b == 0) out[0] = total / kIters;
        if(b == G - 1) out[1] = total / kIters;
        out[3] = (long long)check;
    }
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int G = prop.multiProcessorCount;
    unsigned long long * acc, * sent;
    long long * out;
    const size_t acc_words = (size_t)kIters * kQ * 16, sent_words = (size_t)kIters * kRep * 16;
    cudaMalloc(&acc, acc_words * 8);
    cudaMalloc(&sent, sent_words * 8);
    cudaMalloc(&out, 32);
    for(int work : {0, 3000})
        for(int stride : {16, 4, 1})
        {
            long long h[4] = {0, 0, 0, 0};
            for(int rep = 0; rep < 3; rep++)
            {
                cudaMemset(acc, 0, acc_words * 8);
                cudaMemset(sent, 0, sent_words * 8);
                cudaMemset(out, 0, 32);
                void * args[] = {&acc, &sent, &stride, &out, &work};
                cudaLaunchCooperativeKernel((const void *)probe, dim3(G), dim3(256), args, 0, 0);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
            printf("all-reduce of %d u64 words over %d CTAs, one accumulator per %3d bytes, %4d cycles of work between rounds: %5lld cycles per round (CTA 0), %5lld (last CTA); "
                   "incomplete-at-sentinel reads: %lld\n", kQ, G, stride * 8, work, h[0], h[1], h[2]);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
