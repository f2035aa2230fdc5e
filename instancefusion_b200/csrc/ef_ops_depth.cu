// ef_ops_depth.cu -- the step BEFORE the tracker: ElasticFusion::filterDepth / metriciseDepth, which the
// reference runs as GLSL fragment shaders (elasticfusionpublic/Core/src/Shaders/depth_bilateral.frag:30-76,
// depth_metric.frag:28-40; ElasticFusion.cpp:765-784) between the raw sensor depth and RGBDOdometry::initICP.
// As CUDA they give a GL-free raw-depth -> pose pipeline and remove a GL -> CUDA synchronisation.
//
//   k_depth_bilateral  13x13 bilateral filter of the raw depth (millimetres), range gate [300 mm, maxD], one exp per
//                      tap exactly like the shader: w = exp(-(|dp|^2 / (2 sigma_s^2) + (dv)^2 / (2 sigma_c^2))).
//                      A 32x8 tile + 6-pixel halo is staged once in shared memory as float (3.6 KB); the window clips at the
//                      image border like the shader's loop bounds.  169 exp per pixel: the kernel is bound by the
//                      SFU / FP32 issue rate, not by memory (0.6 MB in, 0.6 MB out at 640x480).
//   k_depth_metric     u16 millimetres -> float metres with the same gate.
//
// Arithmetic is pinned with explicit round-to-nearest intrinsics (no FMA contraction, IEEE division) so that the only
// difference to the CPU restatement the tests check against is the exp implementation (<= 2 ulp): the rounded millimetre
// output agrees except at exact .5 ties.
#include "ef_kernels.h"

namespace ef
{

namespace
{

constexpr int kBfR = 6;                 // depth_bilateral.frag:45
constexpr int kBfTileW = 32, kBfTileH = 8;
constexpr int kBfW = kBfTileW + 2 * kBfR, kBfH = kBfTileH + 2 * kBfR;

// one tap of depth_bilateral.frag:55-67; dx2 + dy2 and its product with sigma_space2_inv_half are compile-time constants
// on the unrolled interior path (the product is folded with the same round-to-nearest multiply)
__device__ __forceinline__ void bf_tap(float tmp, float fv, float space_term, float & sum1, float & sum2)
{
    const float sigma_color2_inv_half = 0.000555556f; // :43
    const float dc = __fsub_rn(fv, tmp);
    const float color2 = __fmul_rn(dc, dc);                                                           // :62
    const float arg = __fadd_rn(space_term, __fmul_rn(color2, sigma_color2_inv_half));
    const float weight = expf(-arg);                                                                  // :64
    sum1 = __fadd_rn(sum1, __fmul_rn(tmp, weight));                                                   // :66
    sum2 = __fadd_rn(sum2, weight);
}

__global__ void __launch_bounds__(kBfTileW * kBfTileH) k_depth_bilateral(const uint16_t * __restrict__ src, int spitch /*elements*/, int rows,
                                                                          int cols, unsigned max_mm, uint16_t * __restrict__ dst,
                                                                          int dpitch /*elements*/)
{
    // the tile is kept as float: every texel is converted once instead of once per tap (13 x 13 reuse)
    __shared__ float tile[kBfH][kBfW + 1];
    const int ox = blockIdx.x * kBfTileW - kBfR, oy = blockIdx.y * kBfTileH - kBfR;
    const int tid = threadIdx.y * kBfTileW + threadIdx.x;
    for(int i = tid; i < kBfW * kBfH; i += kBfTileW * kBfTileH)
    {
        const int ty = i / kBfW, tx = i - ty * kBfW;
        const int gx = ox + tx, gy = oy + ty;
        tile[ty][tx] = (gx >= 0 && gy >= 0 && gx < cols && gy < rows) ? (float)__ldg(src + (size_t)gy * spitch + gx) : 0.f;
    }
    __syncthreads();
    const int x = blockIdx.x * kBfTileW + threadIdx.x, y = blockIdx.y * kBfTileH + threadIdx.y;
    if(x >= cols || y >= rows) return;
    const float fv = tile[threadIdx.y + kBfR][threadIdx.x + kBfR];
    const unsigned value = (unsigned)fv;
    unsigned out = 0;
    if(!(value > max_mm || value < 300u)) // :36
    {
        const float sigma_space2_inv_half = 0.024691358f; // :42
        float sum1 = 0.f, sum2 = 0.f;
        if(x >= kBfR && y >= kBfR && x + kBfR + 1 <= cols && y + kBfR + 1 <= rows)
        {
            // interior: the full 13 x 13 window in the shader's raster order, offsets known at compile time
#pragma unroll
            for(int j = -kBfR; j <= kBfR; j++)
            {
#pragma unroll
                for(int i = -kBfR; i <= kBfR; i++)
                {
                    const float space2 = __fadd_rn(__fmul_rn((float)-i, (float)-i), __fmul_rn((float)-j, (float)-j)); // :61, folded
                    bf_tap(tile[threadIdx.y + kBfR + j][threadIdx.x + kBfR + i], fv, __fmul_rn(space2, sigma_space2_inv_half), sum1, sum2);
                }
            }
        }
        else
        {
            const int x0 = max(x - kBfR, 0), y0 = max(y - kBfR, 0);
            const int tx = min(x + kBfR + 1, cols), ty = min(y + kBfR + 1, rows); // :48-49
            for(int cy = y0; cy < ty; ++cy)
            {
                const float dy = (float)y - (float)cy;
                const float dy2 = __fmul_rn(dy, dy);
                const float * row = tile[cy - oy];
                for(int cx = x0; cx < tx; ++cx)
                {
                    const float dx = (float)x - (float)cx;
                    const float space2 = __fadd_rn(__fmul_rn(dx, dx), dy2);                               // :61
                    bf_tap(row[cx - ox], fv, __fmul_rn(space2, sigma_space2_inv_half), sum1, sum2);
                }
            }
        }
        out = (unsigned)roundf(__fdiv_rn(sum1, sum2)); // :71
    }
    dst[(size_t)y * dpitch + x] = (uint16_t)out;
}

__global__ void __launch_bounds__(256) k_depth_metric(const uint16_t * __restrict__ src, int spitch, int rows, int cols, unsigned max_mm,
                                                      float * __restrict__ dst, int dpitch)
{
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if(x >= cols || y >= rows) return;
    const unsigned value = __ldg(src + (size_t)y * spitch + x);
    dst[(size_t)y * dpitch + x] = (value > max_mm || value < 300u) ? 0.f : __fdiv_rn((float)value, 1000.0f); // depth_metric.frag:30-39
}

} // namespace

cudaError_t launch_depth_bilateral(const uint16_t * src, size_t sp, int rows, int cols, float max_depth_m, uint16_t * dst, size_t dp, cudaStream_t s)
{
    const int spitch = (int)((sp ? sp : (size_t)cols * 2) / 2), dpitch = (int)((dp ? dp : (size_t)cols * 2) / 2);
    const unsigned max_mm = (unsigned)(max_depth_m * 1000.0f); // uint(maxD * 1000.0f)
    const dim3 block(kBfTileW, kBfTileH);
    const dim3 grid((cols + kBfTileW - 1) / kBfTileW, (rows + kBfTileH - 1) / kBfTileH);
    k_depth_bilateral<<<grid, block, 0, s>>>(src, spitch, rows, cols, max_mm, dst, dpitch);
    return cudaGetLastError();
}

cudaError_t launch_depth_metric(const uint16_t * src, size_t sp, int rows, int cols, float max_depth_m, float * dst, size_t dp, cudaStream_t s)
{
    const int spitch = (int)((sp ? sp : (size_t)cols * 2) / 2), dpitch = (int)((dp ? dp : (size_t)cols * 4) / 4);
    const unsigned max_mm = (unsigned)(max_depth_m * 1000.0f);
    const dim3 grid((cols + 63) / 64, (rows + 3) / 4);
    k_depth_metric<<<grid, 256, 0, s>>>(src, spitch, rows, cols, max_mm, dst, dpitch);
    return cudaGetLastError();
}

} // namespace ef
