// ffma2_tput.cu -- microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.ftz.f32x2) on one SM,
// alone and mixed with integer work, 8 warps per SM (the tracker kernel's shape).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_tput tools/ffma2_tput.cu && tools/ffma2_tput
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\tfma.rn.ftz.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
                 : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

constexpr int N = 16; // independent accumulators per thread

template<int MODE> __global__ void __launch_bounds__(256, 1) k(float * out, long long * cycles, int iters, float s)
{
    float a[2 * N];
    int q[N];
#pragma unroll
    for(int i = 0; i < 2 * N; i++) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for(int i = 0; i < N; i++) q[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for(int it = 0; it < iters; it++)
    {
        if(MODE == 0) // 2N scalar FFMA
        {
#pragma unroll
            for(int i = 0; i < 2 * N; i++) a[i] = fmaf(a[i], s, 1.0f);
        }
        else if(MODE == 1) // N packed FFMA2 (the same 2N fused multiply-adds)
        {
#pragma unroll
            for(int i = 0; i < N; i++)
            {
                float2 r = fma2(make_float2(a[2 * i], a[2 * i + 1]), make_float2(s, s), make_float2(1.0f, 1.0f));
                a[2 * i] = r.x;
                a[2 * i + 1] = r.y;
            }
        }
        else if(MODE == 2) // 2N scalar FFMA + N integer multiply-adds
        {
#pragma unroll
            for(int i = 0; i < 2 * N; i++) a[i] = fmaf(a[i], s, 1.0f);
#pragma unroll
            for(int i = 0; i < N; i++) q[i] = q[i] * 3 + it;
        }
        else // N FFMA2 + N integer multiply-adds
        {
#pragma unroll
            for(int i = 0; i < N; i++)
            {
                float2 r = fma2(make_float2(a[2 * i], a[2 * i + 1]), make_float2(s, s), make_float2(1.0f, 1.0f));
                a[2 * i] = r.x;
                a[2 * i + 1] = r.y;
            }
#pragma unroll
            for(int i = 0; i < N; i++) q[i] = q[i] * 3 + it;
        }
    }
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for(int i = 0; i < 2 * N; i++) acc += a[i];
#pragma unroll
    for(int i = 0; i < N; i++) acc += (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main()
{
    float * out;
    long long * cyc;
    cudaMalloc(&out, 148 * 256 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char * names[4] = {"32 FFMA", "16 FFMA2 (= 32 FMA)", "32 FFMA + 16 IMAD", "16 FFMA2 + 16 IMAD"};
    for(int m = 0; m < 4; m++)
    {
        for(int rep = 0; rep < 2; rep++)
        {
            if(m == 0) k<0><<<148, 256>>>(out, cyc, iters, 0.999f);
            if(m == 1) k<1><<<148, 256>>>(out, cyc, iters, 0.999f);
            if(m == 2) k<2><<<148, 256>>>(out, cyc, iters, 0.999f);
            if(m == 3) k<3><<<148, 256>>>(out, cyc, iters, 0.999f);
            cudaDeviceSynchronize();
        }
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-24s %7.2f cycles per loop body per SM (8 warps): %.2f warp-instructions issued per cycle per SM\n", names[m], (double)h / iters,
               8.0 * (m == 0 ? 32 : m == 1 ? 16 : m == 2 ? 48 : 32) / ((double)h / iters));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
