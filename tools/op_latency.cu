// op_latency.cu -- dependent-issue latency (cycles) of the instructions on the tracker's serial path, one warp.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double * out, long long * t, double seed, int one)
{
    const int lane = threadIdx.x;
    double d = seed + lane;
    float f = (float)seed + lane;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 64
    for(int i = 0; i < N; i++) d = fma(d, 1.0000001, 0.5);
    t1 = clock64();
    if(lane == 0) t[0] = t1 - t0;
    // FFMA chain
    t0 = clock64();
#pragma unroll 64
    for(int i = 0; i < N; i++) f = fmaf(f, 1.0000001f, 0.5f);
    t1 = clock64();
    if(lane == 0) t[1] = t1 - t0;
    // SHFL chain (32-bit)
    t0 = clock64();
#pragma unroll 64
    for(int i = 0; i < N; i++) f = __shfl_sync(0xffffffffu, f, (lane + one) & 31);
    t1 = clock64();
    if(lane == 0) t[2] = t1 - t0;
    // double shuffle chain
    t0 = clock64();
#pragma unroll 64
    for(int i = 0; i < N; i++) d = __shfl_sync(0xffffffffu, d, (lane + one) & 31);
    t1 = clock64();
    if(lane == 0) t[3] = t1 - t0;
    // double reciprocal chain
    t0 = clock64();
#pragma unroll 16
    for(int i = 0; i < N; i++) d = 1.0 / d;
    t1 = clock64();
    if(lane == 0) t[4] = t1 - t0;
    // double sqrt chain
    d = fabs(d) + 2.0;
    t0 = clock64();
#pragma unroll 16
    for(int i = 0; i < N; i++) d = sqrt(d) + 1.5;
    t1 = clock64();
    if(lane == 0) t[5] = t1 - t0;
    // sincos
    t0 = clock64();
#pragma unroll 4
    for(int i = 0; i < N; i++) { double s, c; sincos(d * 1e-3, &s, &c); d = s + c; }
    t1 = clock64();
    if(lane == 0) t[6] = t1 - t0;
    // LDS chain (pointer chasing in shared memory)
    __shared__ int sm[64];
    sm[lane] = (lane + one) & 31;
    __syncwarp();
    int p = lane;
    t0 = clock64();
#pragma unroll 64
    for(int i = 0; i < N; i++) p = sm[p];
    t1 = clock64();
    if(lane == 0) t[7] = t1 - t0;
    // F2F f64<->f32 chain
    t0 = clock64();
#pragma unroll 64
    for(int i = 0; i < N; i++) { f = (float)d; d = (double)f + 1.0; }
    t1 = clock64();
    if(lane == 0) t[8] = t1 - t0;
    out[lane] = d + f + p;
}
int main()
{
    double * out; long long * t;
    cudaMalloc(&out, 32 * 8);
    cudaMallocManaged(&t, 16 * 8);
    for(int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, t, 1.25, 1); cudaDeviceSynchronize(); }
    const char * names[] = {"DFMA", "FFMA", "SHFL.32", "shfl double (2 SHFL)", "1.0/d (double)", "sqrt(d)+DADD", "sincos(double)+DADD", "LDS", "F2F f64->f32 + f32->f64 + DADD"};
    for(int i = 0; i < 9; i++) printf("%-34s %.1f cycles\n", names[i], (double)t[i] / N);
    return 0;
}
