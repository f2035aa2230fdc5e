// ef_image_px.cuh -- per-pixel bodies of the image / pyramid operators, shared by the stand-alone
// operator kernels (ef_ops_image.cu) and the fused pyramid builders (ef_build_fused.cu) so that both
// produce the same bits.  Citations: elasticfusionpublic/Core/src/Cuda/cudafuncs.cu.
// Pitches are in ELEMENTS of the respective type.
#pragma once

#include "ef_math.cuh"

namespace ef
{

// pyrDownGaussKernel  cudafuncs.cu:57-94: output pixel (x, y) of the half-resolution u16 depth
__device__ __forceinline__ uint16_t pyr_down_u16_px(const uint16_t * __restrict__ src, int sp, int srows, int scols, int x, int y)
{
    const int D = 5;
    const float sigma_color = 30.f; // :103
    const int center = __ldg(src + (size_t)(2 * y) * sp + 2 * x);

    const int x_mi = max(0, 2 * x - D / 2) - 2 * x;
    const int y_mi = max(0, 2 * y - D / 2) - 2 * y;
    const int x_ma = min(scols, 2 * x - D / 2 + D) - 2 * x;
    const int y_ma = min(srows, 2 * y - D / 2 + D) - 2 * y;

    float sum = 0;
    float wall = 0;
    const float weights[] = {0.375f, 0.25f, 0.0625f};

    for(int yi = y_mi; yi < y_ma; ++yi)
        for(int xi = x_mi; xi < x_ma; ++xi)
        {
            const int val = __ldg(src + (size_t)(2 * y + yi) * sp + 2 * x + xi);
            if(abs(val - center) < 3 * sigma_color)
            {
                sum += val * weights[abs(xi)] * weights[abs(yi)];
                wall += weights[abs(xi)] * weights[abs(yi)];
            }
        }
    return static_cast<int>(sum / wall);
}

// computeVmapKernel  cudafuncs.cu:116-131: vertex of pixel (u, v) from raw depth; returns false (invalid) when
// the depth is 0 or beyond the cutoff
__device__ __forceinline__ bool vertex_px(uint16_t d, int u, int v, float fx_inv, float fy_inv, float cx, float cy, float cutoff, float3 & vtx)
{
    // `depth / 1000.f` compiles to a multiply by 0x3A83126F under --prec-div=false (reference PTX).  Every product
    // is rounded explicitly (mul.rn cannot be contracted): when a fused kernel consumes the vertex in the same
    // thread (normals = differences of vertices) nvcc would otherwise fold these multiplies into the consumer's
    // FMAs and skip the rounding the stored vertex map has.
    const float z = __fmul_rn((float)d, 0.001f);
    if(z != 0 && z < cutoff)
    {
        vtx.x = __fmul_rn(__fmul_rn(z, (u - cx)), fx_inv);
        vtx.y = __fmul_rn(__fmul_rn(z, (v - cy)), fy_inv);
        vtx.z = z;
        return true;
    }
    return false;
}

// computeNmapKernel  cudafuncs.cu:180
__device__ __forceinline__ float3 normal_px(const float3 & v00, const float3 & v01, const float3 & v10)
{
    return normalized3(cross3(v01 - v00, v10 - v00));
}

// {1,4,6,4,1} (x) {1,4,6,4,1}, i = r*5 + c   (cudafuncs.cu:453-457)
__device__ __forceinline__ float gauss5(int i)
{
    const float k1[5] = {1.f, 4.f, 6.f, 4.f, 1.f};
    return k1[i / 5] * k1[i % 5];
}

// pyrDownKernelGaussF  cudafuncs.cu:332-363, reading the source through `at(cy, cx)`
template<class F>
__device__ __forceinline__ float pyr_down_gauss_f32_px(F at, int srows, int scols, int x, int y)
{
    const int D = 5;
    if(x >= 1 && y >= 1 && 2 * x + 3 <= scols - 1 && 2 * y + 3 <= srows - 1)
    {
        // interior: no clamp is active, the window is the full 5x5 and the (mirrored) kernel index runs 24..0;
        // same raster order and the same FMA chain as the general loop, with the binomial taps as literals
        const float k1[5] = {1.f, 4.f, 6.f, 4.f, 1.f};
        float sum = 0;
        int count = 0;
#pragma unroll
        for(int j = 0; j < 5; j++)
#pragma unroll
            for(int i = 0; i < 5; i++)
            {
                const float s = at(2 * y - 2 + j, 2 * x - 2 + i);
                if(!isnan(s))
                {
                    const float k = k1[4 - j] * k1[4 - i];
                    sum = __fmaf_rn(s, k, sum);
                    count += k;
                }
            }
        return (float)(sum / (float)count);
    }
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int cy = max(0, 2 * y - D / 2);
    float sum = 0;
    int count = 0;
    for(; cy < ty; ++cy)
        for(int cx = max(0, 2 * x - D / 2); cx < tx; ++cx)
        {
            const float s = at(cy, cx);
            if(!isnan(s))
            {
                const float k = gauss5((ty - cy - 1) * 5 + (tx - cx - 1));
                sum = __fmaf_rn(s, k, sum); // the reference binary contracts `sum += s * k` into one FMA
                count += k;
            }
        }
    return (float)(sum / (float)count);
}

// pyrDownKernelIntensityGauss  cudafuncs.cu:470-500
template<class F>
__device__ __forceinline__ uint8_t pyr_down_gauss_u8_px(F at, int srows, int scols, int x, int y)
{
    const int D = 5;
    if(x >= 1 && y >= 1 && 2 * x + 3 <= scols - 1 && 2 * y + 3 <= srows - 1)
    {
        // interior fast path (all sums are exact small integers in binary32, so only the taps' membership matters)
        const int k1[5] = {1, 4, 6, 4, 1};
        int isum = 0, count = 0;
#pragma unroll
        for(int j = 0; j < 5; j++)
#pragma unroll
            for(int i = 0; i < 5; i++)
            {
                const int s = at(2 * y - 2 + j, 2 * x - 2 + i);
                const int k = k1[j] * k1[i];
                if(s > 0)
                {
                    isum += s * k;
                    count += k;
                }
            }
        const uint8_t r = ((float)isum / (float)count);
        return r;
    }
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int cy = max(0, 2 * y - D / 2);
    float sum = 0;
    int count = 0;
    for(; cy < ty; ++cy)
        for(int cx = max(0, 2 * x - D / 2); cx < tx; ++cx)
        {
            const uint8_t s = at(cy, cx);
            if(s > 0)
            {
                const float k = gauss5((ty - cy - 1) * 5 + (tx - cx - 1));
                sum += s * k;
                count += k;
            }
        }
    const uint8_t r = (sum / (float)count);
    return r;
}

// bgr2IntensityKernel  cudafuncs.cu:560
__device__ __forceinline__ uint8_t intensity_px(uchar4 s)
{
    // x*0.114f + y*0.299f + z*0.587f as the reference binary contracts it: t = y*.299; t = fma(x, .114, t); t = fma(z, .587, t)
    const int value = __fmaf_rn((float)s.z, 0.587f, __fmaf_rn((float)s.x, 0.114f, (float)s.y * 0.299f));
    return (uint8_t)value;
}

// verticesToDepthKernel  cudafuncs.cu:536
__device__ __forceinline__ float depth_from_z(float z, float cutoff) { return (z > cutoff || z <= 0) ? qnan() : z; }

// applyKernel taps (cudafuncs.cu:615-621)
__device__ __forceinline__ float sobel_x_tap(int k)
{
    const float t[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    return t[k];
}
__device__ __forceinline__ float sobel_y_tap(int k)
{
    const float t[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
    return t[k];
}

// applyKernel  cudafuncs.cu:583-607: `kernelIndex` counts down over VISITED taps only (:594-603)
template<class F>
__device__ __forceinline__ void derivative_px(F at, int rows, int cols, int x, int y, short & dx, short & dy)
{
    if(x >= 1 && y >= 1 && x + 1 <= cols - 1 && y + 1 <= rows - 1)
    {
        // interior: all nine taps are visited, kernelIndex runs 8..0 in raster order.  Zero taps are skipped
        // (fma(s, +-0, acc) == acc exactly); the non-zero ones keep their order.
        const float tl = (float)at(y - 1, x - 1), tm = (float)at(y - 1, x), tr = (float)at(y - 1, x + 1);
        const float ml = (float)at(y, x - 1), mr = (float)at(y, x + 1);
        const float bl = (float)at(y + 1, x - 1), bm = (float)at(y + 1, x), br = (float)at(y + 1, x + 1);
        float gx = 0, gy = 0;
        gx = __fmaf_rn(tl, -0.52201f, gx);
        gx = __fmaf_rn(tr, 0.52201f, gx);
        gx = __fmaf_rn(ml, -0.79451f, gx);
        gx = __fmaf_rn(mr, 0.79451f, gx);
        gx = __fmaf_rn(bl, -0.52201f, gx);
        gx = __fmaf_rn(br, 0.52201f, gx);
        gy = __fmaf_rn(tl, -0.52201f, gy);
        gy = __fmaf_rn(tm, -0.79451f, gy);
        gy = __fmaf_rn(tr, -0.52201f, gy);
        gy = __fmaf_rn(bl, 0.52201f, gy);
        gy = __fmaf_rn(bm, 0.79451f, gy);
        gy = __fmaf_rn(br, 0.52201f, gy);
        dx = gx;
        dy = gy;
        return;
    }
    float dxVal = 0;
    float dyVal = 0;
    int kernelIndex = 8;
    for(int j = max(y - 1, 0); j <= min(y + 1, rows - 1); j++)
        for(int i = max(x - 1, 0); i <= min(x + 1, cols - 1); i++)
        {
            const float s = (float)at(j, i);
            dxVal = __fmaf_rn(s, sobel_x_tap(kernelIndex), dxVal);
            dyVal = __fmaf_rn(s, sobel_y_tap(kernelIndex), dyVal);
            --kernelIndex;
        }
    dx = dxVal;
    dy = dyVal;
}

} // namespace ef
