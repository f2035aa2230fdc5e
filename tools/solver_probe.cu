// solver_probe.cu -- microbenchmark of the 6x6 normal-equation solve of the tracker kernel's solver warp: cycles of one warp
// (every lane the same scalar routine) for  (a) LDL^T with the Newton reciprocal on the pivot chain (the product),
// (b) division-free elimination with the reciprocals off the chain,  (c) the input stage alone (convert 27 floats, park
// them in shared memory, read them back as broadcast 128-bit loads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I instancefusion_b200/csrc -o tools/solver_probe tools/solver_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ef_hostmath.h"
using namespace ef;

__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return (d != 0.0) ? r : 0.0;
}

template<int MODE> __global__ void probe(const float * in, double * out, long long * cyc, int reps)
{
    __shared__ double s_S[28];
    __shared__ float s_final[64];
    const int lane = threadIdx.x & 31;
    if(threadIdx.x < 58) s_final[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    double acc = 0;
    const long long t0 = clock64();
    for(int rep = 0; rep < reps; rep++)
    {
        const int sl = lane < 27 ? lane : 27;
        s_S[sl] = (double)s_final[sl] * 100.0 + (double)s_final[29 + sl] + acc * 1e-30;
        __syncwarp();
        double a[28];
#pragma unroll
        for(int k = 0; k < 14; k++)
        {
            const double2 v = reinterpret_cast<const double2 *>(s_S)[k];
            a[2 * k] = v.x;
            a[2 * k + 1] = v.y;
        }
        double x[6];
        if(MODE == 2)
        {
#pragma unroll
            for(int i = 0; i < 6; i++) x[i] = a[i] + a[6 + i];
        }
        else if(MODE == 0)
        {
            double inv[6];
#pragma unroll
            for(int j = 0; j < 6; j++)
            {
                inv[j] = fast_rcp(a[hm::acc_index(j, j)]);
                double l[6];
#pragma unroll
                for(int i = j + 1; i < 6; i++) l[i] = a[hm::acc_index(j, i)] * inv[j];
#pragma unroll
                for(int i = j + 1; i < 6; i++)
#pragma unroll
                    for(int k = i; k < 7; k++) a[hm::acc_index(i, k)] = fma(-l[i], a[hm::acc_index(j, k)], a[hm::acc_index(i, k)]);
#pragma unroll
                for(int i = j + 1; i < 6; i++) a[hm::acc_index(j, i)] = l[i];
            }
            double wv[6];
#pragma unroll
            for(int i = 0; i < 6; i++) wv[i] = a[hm::acc_index(i, 6)] * inv[i];
#pragma unroll
            for(int i = 5; i >= 0; i--)
            {
                x[i] = wv[i];
#pragma unroll
                for(int r = 0; r < i; r++) wv[r] = fma(-a[hm::acc_index(r, i)], x[i], wv[r]);
            }
        }
        else
        {
            double inv[6];
#pragma unroll
            for(int j = 0; j < 6; j++)
            {
                const double p = a[hm::acc_index(j, j)];
                inv[j] = fast_rcp(p);
                if(j < 5)
                {
                    const double s2 = __hiloint2double(0x7fe00000 - (__double2hiint(p) & 0x7ff00000), 0);
                    const double m = p * s2;
                    double l[6];
#pragma unroll
                    for(int i = j + 1; i < 6; i++) l[i] = a[hm::acc_index(j, i)] * s2;
#pragma unroll
                    for(int i = j + 1; i < 6; i++)
#pragma unroll
                        for(int k = i; k < 7; k++) a[hm::acc_index(i, k)] = fma(-l[i], a[hm::acc_index(j, k)], m * a[hm::acc_index(i, k)]);
                }
            }
            double wv[6];
#pragma unroll
            for(int i = 0; i < 6; i++) wv[i] = a[hm::acc_index(i, 6)];
#pragma unroll
            for(int i = 5; i >= 0; i--)
            {
                x[i] = wv[i] * inv[i];
#pragma unroll
                for(int r = 0; r < i; r++) wv[r] = fma(-a[hm::acc_index(r, i)], x[i], wv[r]);
            }
        }
        acc = x[0] + x[1] + x[2] + x[3] + x[4] + x[5];
        __syncwarp();
    }
    const long long t1 = clock64();
    if(threadIdx.x == 0)
    {
        *cyc = (t1 - t0) / reps;
        *out = acc;
    }
}

int main()
{
    // an SPD system: A = M^T M + I packed as the tracker's accumulator (27 upper-triangle entries incl. the right-hand side)
    float h[58] = {0};
    double M[6][7];
    unsigned s = 12345;
    for(int i = 0; i < 6; i++)
        for(int j = 0; j < 7; j++) { s = s * 1664525u + 1013904223u; M[i][j] = ((s >> 8) % 2000) / 1000.0 - 1.0; }
    int k = 0;
    for(int i = 0; i < 6; i++)
        for(int j = i; j < 7; j++)
        {
            double v = 0;
            for(int r = 0; r < 6; r++) v += M[r][i] * M[r][j];
            if(i == j) v += 1.0;
            h[k] = (float)(v * 1e5);
            h[29 + k] = (float)(v * 3e6);
            k++;
        }
    float * d_in; double * d_out; long long * d_cyc;
    cudaMalloc(&d_in, sizeof(h)); cudaMalloc(&d_out, 8); cudaMalloc(&d_cyc, 8);
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    const char * names[3] = {"LDL^T, Newton reciprocal on the pivot chain (product)", "division-free elimination, reciprocals off the chain", "input stage only (convert, park, broadcast loads)"};
    for(int m = 0; m < 3; m++)
    {
        for(int rep = 0; rep < 2; rep++)
        {
            if(m == 0) probe<0><<<1, 32>>>(d_in, d_out, d_cyc, 1000);
            if(m == 1) probe<1><<<1, 32>>>(d_in, d_out, d_cyc, 1000);
            if(m == 2) probe<2><<<1, 32>>>(d_in, d_out, d_cyc, 1000);
            cudaDeviceSynchronize();
        }
        long long c; double o;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&o, d_out, 8, cudaMemcpyDeviceToHost);
        printf("%-60s %5lld cycles per solve (sum x = %.12g)\n", names[m], c, o);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
