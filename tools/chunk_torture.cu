// chunk_torture.cu -- torture test of the flagged-chunk hand-off of ef_track_kernel.cu: payload and flag travel in ONE
// 128-bit relaxed store and are observed by ONE 128-bit relaxed load.  A torn chunk (words of two different stores in
// one observation) would silently corrupt a pose, so this hammers exactly that:
//
//   every CTA (one per SM, cooperative launch = all co-resident) owns kChunks chunks; its writer threads store
//   {n, n ^ A, n * B, n} for n = 1, 2, 3, ... back to back, never waiting;
//   every other thread of every CTA polls a chunk of ANOTHER CTA (all ordered SM pairs are covered by rotating the
//   partner every kRotate reads) and checks EVERY value it observes: the four words must belong to one store
//   (y == x ^ A, z == x * B, w == x) and, per location, x must never go backwards (coherence);
//   meanwhile a second kernel on another stream streams a large buffer through HBM (copy with a twist), so the
//   chunk traffic shares L2 and the memory system with ordinary loads and stores.
//
// Build / run (on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chunk_torture tools/chunk_torture.cu
//   tools/chunk_torture [seconds=10] [mode: 0 = .b128 scalar forms (the product), 1 = .v4.u32 vector forms]
// Prints: observations, distinct new values observed (= exchanges), torn chunks, backward steps.  Exit code 1 on any.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr unsigned kA = 0x5bd1e995u, kB = 2654435761u;
constexpr int kChunks = 8;      // chunks per CTA (the parameter line of the tracker has 8)
constexpr int kThreads = 256;
constexpr int kRotate = 64;     // reads of one partner before moving to the next

template<int MODE> __device__ __forceinline__ uint4 ld_chunk(const uint4 * p)
{
    uint4 v;
    if(MODE == 0)
        asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%4];\n\tmov.b128 {%0, %1, %2, %3}, q;\n\t}"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template<int MODE> __device__ __forceinline__ void st_chunk(uint4 * p, uint4 v)
{
    if(MODE == 0)
        asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p), "r"(v.x), "r"(v.y),
                     "r"(v.z), "r"(v.w) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct Counters
{
    unsigned long long reads, fresh, torn, backward;
};

template<int MODE> __global__ void __launch_bounds__(kThreads, 1) torture(uint4 * chunks, volatile int * stop, Counters * out, long long max_cycles)
{
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    const long long t0 = clock64();
    unsigned long long reads = 0, fresh = 0, torn = 0, backward = 0;
    if(t < kChunks)
    {
        uint4 * mine = chunks + (size_t)b * kChunks + t;
        unsigned n = 0;
        while(true)
        {
            ++n;
            st_chunk<MODE>(mine, make_uint4(n, n ^ kA, n * kB, n));
            if((n & 255u) == 0 && (clock64() - t0 > max_cycles || *stop)) break;
        }
        fresh = n; // stores made (reported separately below)
    }
    else if(t >= 32) // (lanes 8..31 of the writers' warp stay out of it)
    {
        const int r = t - 32;
        int hop = 1 + r % (G - 1);
        unsigned last = 0;
        const uint4 * p = chunks + (size_t)((b + hop) % G) * kChunks + (r % kChunks);
        int left = kRotate;
        while(true)
        {
            const uint4 v = ld_chunk<MODE>(p);
            ++reads;
            if(v.x != 0u || v.w != 0u)
            {
                if(v.y != (v.x ^ kA) || v.z != v.x * kB || v.w != v.x) ++torn;
                if((int)(v.x - last) < 0) ++backward;
                if(v.x != last) ++fresh;
                last = v.x;
            }
            if(--left == 0)
            {
                left = kRotate;
                hop = 1 + (hop % (G - 1)); // next partner: over time every (reader SM, writer SM) pair is visited
                p = chunks + (size_t)((b + hop) % G) * kChunks + (r % kChunks);
                last = 0;
                if(clock64() - t0 > max_cycles || *stop) break;
            }
        }
    }
    if(t >= 32)
    {
        atomicAdd(&out->reads, reads);
        atomicAdd(&out->fresh, fresh);
        atomicAdd(&out->torn, torn);
        atomicAdd(&out->backward, backward);
    }
    else if(t < kChunks)
        atomicAdd(&out[1].fresh, fresh); // stores
}

// background HBM traffic: y = x + 1 over a buffer far larger than L2, until told to stop
__global__ void churn(const uint4 * __restrict__ x, uint4 * __restrict__ y, size_t n, volatile int * stop)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    while(!*stop)
        for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        {
            uint4 v = x[i];
            v.x += 1;
            y[i] = v;
        }
}

#define CK(x)                                                                      \
    do                                                                             \
    {                                                                              \
        cudaError_t e = (x);                                                       \
        if(e != cudaSuccess)                                                       \
        {                                                                          \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                \
            return 2;                                                              \
        }                                                                          \
    } while(0)

int main(int argc, char ** argv)
{
    const double seconds = argc > 1 ? atof(argv[1]) : 10.0;
    const int mode = argc > 2 ? atoi(argv[2]) : 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    // the torture kernel takes every SM but 8; the churn kernel gets the rest (both must be resident together)
    const int G = sms - 8;
    uint4 * chunks;
    Counters * out;
    int * stop;
    CK(cudaMalloc(&chunks, (size_t)G * kChunks * sizeof(uint4)));
    CK(cudaMemset(chunks, 0, (size_t)G * kChunks * sizeof(uint4)));
    CK(cudaMalloc(&out, 2 * sizeof(Counters)));
    CK(cudaMemset(out, 0, 2 * sizeof(Counters)));
    CK(cudaHostAlloc((void **)&stop, sizeof(int), cudaHostAllocMapped));
    *stop = 0;
    const size_t n = (size_t)1 << 26; // 2 x 1 GiB
    uint4 * x, * y;
    CK(cudaMalloc(&x, n * sizeof(uint4)));
    CK(cudaMalloc(&y, n * sizeof(uint4)));
    CK(cudaMemset(x, 1, n * sizeof(uint4)));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    long long max_cycles = (long long)(seconds * khz * 1e3);
    int * dstop = stop;
    void * args[] = {&chunks, &dstop, &out, &max_cycles};
    const void * fn = mode == 0 ? (const void *)torture<0> : (const void *)torture<1>;
    CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kThreads), args, 0, s1));
    churn<<<8, 1024, 0, s2>>>(x, y, n, dstop);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s1));
    *stop = 1;
    CK(cudaStreamSynchronize(s2));
    Counters h[2];
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("{\"tool\": \"chunk_torture\", \"gpu\": \"%s\", \"access\": \"%s\", \"seconds\": %.1f, \"ctas\": %d, \"reader_threads\": %d, "
           "\"stores\": %llu, \"observations\": %llu, \"exchanges_observed\": %llu, \"torn\": %llu, \"backward\": %llu}\n",
           prop.name, mode == 0 ? "ld/st.relaxed.gpu.global.b128" : "ld/st.relaxed.gpu.global.v4.u32", seconds, G, G * (kThreads - 32),
           h[1].fresh, h[0].reads, h[0].fresh, h[0].torn, h[0].backward);
    return (h[0].torn || h[0].backward) ? 1 : 0;
}
