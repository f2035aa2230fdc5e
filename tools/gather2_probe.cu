// gather2_probe.cu -- microbenchmark: how long until EVERY CTA knows the sum of all CTAs' 58-float partial rows?
//   A  the tracker kernel's scheme: CTA 0 gathers the 147 rows (20 flagged 16-byte chunks each), adds them, and publishes a
//      24-float parameter line to 32 replicas; the workers poll their replica.            rows -> gather -> [solve] -> broadcast
//   B  two-level all-gather: C collector CTAs add the rows of their group and publish a partial row to R replicas; every CTA
//      polls the C partial rows and adds them itself (and would then run the solve redundantly).  rows -> collect -> all-gather [-> solve]
// Cycles from a CTA's own row publication to the moment it holds the result, averaged over the CTAs and rounds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather2_probe tools/gather2_probe.cu && tools/gather2_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_chunk(const uint4 * p)
{
    uint4 v;
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%4];\n\tmov.b128 {%0, %1, %2, %3}, q;\n\t}" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_chunk(uint4 * p, const uint4 & v)
{
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int kChunks = 20, kFloats = 60, kT = 256;

__global__ void __launch_bounds__(kT, 1) scheme_a(uint4 * rows, uint4 * par, int iters, int skew, long long * out)
{
    __shared__ float s_rows[160 * kFloats];
    __shared__ float s_red[4 * 64], s_fin[64];
    const int G = gridDim.x, W = G - 1, b = blockIdx.x, t = threadIdx.x;
    long long total = 0;
    for(int n = 1; n <= iters; n++)
    {
        const long long w0 = clock64();
        while(clock64() - w0 < 200 + (b % 5) * skew) {}
        __syncthreads();
        const long long t0 = clock64();
        if(b > 0)
        {
            if(t < kChunks) st_chunk(rows + (b - 1) * kChunks + t, make_uint4(b, t, n, n));
            if(t < 8)
            {
                uint4 v;
                do { v = ld_chunk(par + ((b - 1) % 32) * 16 + t); } while(v.w != (unsigned)n);
                s_fin[t] = __uint_as_float(v.x);
            }
            __syncthreads();
        }
        else
        {
            const int totalc = W * kChunks;
            for(int i0 = t; i0 < totalc; i0 += 12 * kT)
            {
                uint4 v[12];
                unsigned todo = 0;
#pragma unroll
                for(int k = 0; k < 12; k++) if(i0 + k * kT < totalc) todo |= 1u << k;
                while(todo)
                {
#pragma unroll
                    for(int k = 0; k < 12; k++) if(todo & (1u << k)) v[k] = ld_chunk(rows + i0 + k * kT);
#pragma unroll
                    for(int k = 0; k < 12; k++)
                        if((todo & (1u << k)) && v[k].w == (unsigned)n)
                        {
                            todo &= ~(1u << k);
                            float * d = s_rows + 3 * (i0 + k * kT);
                            d[0] = __uint_as_float(v[k].x); d[1] = __uint_as_float(v[k].y); d[2] = __uint_as_float(v[k].z);
                        }
                }
            }
            __syncthreads();
            const int slot = t & 63, part = t >> 6;
            float s = 0.f;
            if(slot < kFloats) for(int w = part; w < W; w += 4) s += s_rows[w * kFloats + slot];
            s_red[part * 64 + slot] = s;
            __syncthreads();
            if(t < 64) s_fin[t] = s_red[t] + s_red[64 + t] + s_red[128 + t] + s_red[192 + t];
            __syncthreads();
            // (the solve would run here)
            if(t < 32)
                for(int i = t; i < 32 * 8; i += 32) st_chunk(par + (i / 8) * 16 + (i % 8), make_uint4(__float_as_uint(s_fin[i % 8]), 0, 0, n));
        }
        total += clock64() - t0;
    }
    if(t == 0) out[b] = total;
}

// C collectors (CTAs 0 .. C-1, which are workers too); row r belongs to collector r % C; R replicas of every partial row,
// two copies by round parity (a slow CTA of another group may still be reading round n when a collector publishes n + 1)
__global__ void __launch_bounds__(kT, 1) scheme_b(uint4 * rows, uint4 * part_rows, int C, int R, int iters, int skew, long long * out)
{
    __shared__ float s_rows[40 * kFloats];
    __shared__ float s_part[16 * kFloats];
    __shared__ float s_fin[64];
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    long long total = 0;
    for(int n = 1; n <= iters; n++)
    {
        const long long w0 = clock64();
        while(clock64() - w0 < 200 + (b % 5) * skew) {}
        __syncthreads();
        const long long t0 = clock64();
        if(t < kChunks) st_chunk(rows + b * kChunks + t, make_uint4(b, t, n, n));
        if(b < C)
        {
            // collector: rows b, b + C, b + 2C, ...
            const int mine = (G - b + C - 1) / C, totalc = mine * kChunks;
            for(int i0 = t; i0 < totalc; i0 += 2 * kT)
            {
                uint4 v[2];
                unsigned todo = 1u | ((i0 + kT < totalc) ? 2u : 0u);
                while(todo)
                {
#pragma unroll
                    for(int k = 0; k < 2; k++)
                        if(todo & (1u << k))
                        {
                            const int i = i0 + k * kT, r = i / kChunks, c = i - r * kChunks;
                            v[k] = ld_chunk(rows + (b + r * C) * kChunks + c);
                        }
#pragma unroll
                    for(int k = 0; k < 2; k++)
                        if((todo & (1u << k)) && v[k].w == (unsigned)n)
                        {
                            todo &= ~(1u << k);
                            float * d = s_rows + 3 * (i0 + k * kT);
                            d[0] = __uint_as_float(v[k].x); d[1] = __uint_as_float(v[k].y); d[2] = __uint_as_float(v[k].z);
                        }
                }
            }
            __syncthreads();
            if(t < kFloats)
            {
                float s = 0.f;
                for(int r = 0; r < mine; r++) s += s_rows[r * kFloats + t];
                s_fin[t] = s;
            }
            __syncthreads();
            for(int i = t; i < R * kChunks; i += kT)
            {
                const int rep = i / kChunks, c = i - rep * kChunks;
                st_chunk(part_rows + ((size_t)((n & 1) * R + rep) * C + b) * 32 + c, make_uint4(__float_as_uint(s_fin[3 * c]), __float_as_uint(s_fin[3 * c + 1]), __float_as_uint(s_fin[3 * c + 2]), n));
            }
        }
        // everybody: the C partial rows of replica b % R
        for(int i = t; i < C * kChunks; i += kT)
        {
            const int col = i / kChunks, c = i - col * kChunks;
            uint4 v;
            do { v = ld_chunk(part_rows + ((size_t)((n & 1) * R + b % R) * C + col) * 32 + c); } while(v.w != (unsigned)n);
            float * d = s_part + 3 * i;
            d[0] = __uint_as_float(v.x); d[1] = __uint_as_float(v.y); d[2] = __uint_as_float(v.z);
        }
        __syncthreads();
        if(t < kFloats)
        {
            float s = 0.f;
            for(int col = 0; col < C; col++) s += s_part[col * kFloats + t];
            s_fin[t] = s;
        }
        __syncthreads();
        total += clock64() - t0;
    }
    if(t == 0) out[b] = total;
}

// C  one-hop all-gather: every CTA polls ALL rows itself (12 chunks in flight per thread) and adds them in row order -- the
//    sums, and a redundant solve, are then bit-identical in every CTA.  Rows in two copies by round parity.
__global__ void __launch_bounds__(kT, 1) scheme_c(uint4 * rows, int iters, int skew, int stagger, long long * out)
{
    extern __shared__ float s_all[]; // G * kFloats
    __shared__ float s_red[4 * 64], s_fin[64];
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    long long total = 0;
    for(int n = 1; n <= iters; n++)
    {
        const long long w0 = clock64();
        while(clock64() - w0 < 200 + (b % 5) * skew) {}
        __syncthreads();
        const long long t0 = clock64();
        uint4 * R = rows + (size_t)(n & 1) * 256 * kChunks;
        if(t < kChunks) st_chunk(R + b * kChunks + t, make_uint4(b, t, n, n));
        const int totalc = G * kChunks;
        // (stagger: CTA b starts its sweep at row b, so that the 148 readers do not all hit the same line at the same time)
        const int shift = stagger ? b * kChunks : 0;
        for(int i0 = t; i0 < totalc; i0 += 12 * kT)
        {
            uint4 v[12];
            unsigned todo = 0;
#pragma unroll
            for(int k = 0; k < 12; k++) if(i0 + k * kT < totalc) todo |= 1u << k;
            while(todo)
            {
#pragma unroll
                for(int k = 0; k < 12; k++)
                    if(todo & (1u << k))
                    {
                        int i = i0 + k * kT + shift;
                        if(i >= totalc) i -= totalc;
                        v[k] = ld_chunk(R + i);
                    }
#pragma unroll
                for(int k = 0; k < 12; k++)
                    if((todo & (1u << k)) && v[k].w == (unsigned)n)
                    {
                        todo &= ~(1u << k);
                        int i = i0 + k * kT + shift;
                        if(i >= totalc) i -= totalc;
                        float * d = s_all + 3 * i;
                        d[0] = __uint_as_float(v[k].x); d[1] = __uint_as_float(v[k].y); d[2] = __uint_as_float(v[k].z);
                    }
            }
        }
        __syncthreads();
        const int slot = t & 63, part = t >> 6;
        float s = 0.f;
        if(slot < kFloats) for(int w = part; w < G; w += 4) s += s_all[w * kFloats + slot];
        s_red[part * 64 + slot] = s;
        __syncthreads();
        if(t < 64) s_fin[t] = s_red[t] + s_red[64 + t] + s_red[128 + t] + s_red[192 + t];
        __syncthreads();
        total += clock64() - t0;
    }
    if(t == 0) out[b] = total + (long long)(s_fin[0] == 12345.f);
}

int main()
{
    uint4 * buf;
    long long * out;
    cudaMalloc(&buf, 8 << 20);
    cudaMallocManaged(&out, 256 * sizeof(long long));
    int iters = 2000;
    const int G = 148;
    for(int skew : {0, 100})
    {
        cudaMemset(buf, 0, 8 << 20);
        uint4 * rows = buf, * par = buf + 65536;
        void * a[] = {&rows, &par, &iters, &skew, &out};
        cudaLaunchCooperativeKernel((const void *)scheme_a, dim3(G), dim3(kT), a, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        double w = 0, mx = 0;
        for(int b = 1; b < G; b++) { w += out[b]; if(out[b] > mx) mx = out[b]; }
        printf("A central gather + broadcast (no solve), skew %3d: CTA 0 %6.0f cycles per round, workers mean %6.0f max %6.0f (%s)\n", skew, (double)out[0] / iters,
               w / (G - 1) / iters, mx / iters, cudaGetErrorString(e));
        for(int stagger : {0, 1})
        {
            cudaMemset(buf, 0, 8 << 20);
            void * cc[] = {&rows, &iters, &skew, &stagger, &out};
            cudaFuncSetAttribute(scheme_c, cudaFuncAttributeMaxDynamicSharedMemorySize, 148 * kFloats * 4);
            cudaLaunchCooperativeKernel((const void *)scheme_c, dim3(G), dim3(kT), cc, 148 * kFloats * 4, 0);
            e = cudaDeviceSynchronize();
            w = 0; mx = 0;
            for(int b = 0; b < G; b++) { w += out[b]; if(out[b] > mx) mx = out[b]; }
            printf("C one-hop all-gather, stagger %d, skew %3d: mean %6.0f max %6.0f cycles per round (%s)\n", stagger, skew, w / G / iters, mx / iters, cudaGetErrorString(e));
        }
        for(int C : {8})
            for(int R : {1})
            {
                cudaMemset(buf, 0, 8 << 20);
                void * bb[] = {&rows, &par, &C, &R, &iters, &skew, &out};
                cudaLaunchCooperativeKernel((const void *)scheme_b, dim3(G), dim3(kT), bb, 0, 0);
                e = cudaDeviceSynchronize();
                w = 0; mx = 0;
                for(int b = 0; b < G; b++) { w += out[b]; if(out[b] > mx) mx = out[b]; }
                printf("B %2d collectors x %d replicas, skew %3d: mean %6.0f max %6.0f cycles per round (%s)\n", C, R, skew, w / G / iters, mx / iters, cudaGetErrorString(e));
            }
    }
    return 0;
}
