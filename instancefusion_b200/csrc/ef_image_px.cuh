// ef_image_px.cuh -- per-pixel bodies of the image / pyramid operators, shared by the stand-alone
// operator kernels (ef_ops_image.cu) and the fused pyramid builders (ef_build_fused.cu) so that both
// produce the same bits.  Citations: elasticfusionpublic/Core/src/Cuda/cudafuncs.cu.
// Pitches are in ELEMENTS of the respective type.
#pragma once

#include "ef_math.cuh"

namespace ef
{

// pyrDownGaussKernel  cudafuncs.cu:57-94: output pixel (x, y) of the half-resolution u16 depth, the source read
// through `at(row, col)` (global memory or a shared-memory tile)
// Integer evaluation: the taps are {6, 4, 1} / 16 per axis, so every product val * w(xi) * w(yi) is a multiple of 2^-8 below
// 2^16 * 9/64 and the float sums the reference forms (`sum`, `wall`) are exact in binary32 whatever their order -- the same
// two floats come out of integer accumulators scaled by 2^-8.  Only the final division rounds, and it is the same operation.
template<class F>
__device__ __forceinline__ uint16_t pyr_down_u16_at(F at, int srows, int scols, int x, int y)
{
    const int D = 5;
    const int color_gate = 90; // 3 * sigma_color, sigma_color = 30 (:64, :103)
    const int center = at(2 * y, 2 * x);
    int isum = 0, iwall = 0;
    if(x >= 1 && y >= 1 && 2 * x + 2 < scols && 2 * y + 2 < srows)
    {
        // interior: no clamp is active
#pragma unroll
        for(int yi = -2; yi <= 2; ++yi)
#pragma unroll
            for(int xi = -2; xi <= 2; ++xi)
            {
                const int val = at(2 * y + yi, 2 * x + xi);
                const int w = ((xi == 0) ? 6 : (xi == 1 || xi == -1) ? 4 : 1) * ((yi == 0) ? 6 : (yi == 1 || yi == -1) ? 4 : 1);
                const bool in = abs(val - center) < color_gate;
                isum += in ? val * w : 0;
                iwall += in ? w : 0;
            }
    }
    else
    {
        const int x_mi = max(0, 2 * x - D / 2) - 2 * x;
        const int y_mi = max(0, 2 * y - D / 2) - 2 * y;
        const int x_ma = min(scols, 2 * x - D / 2 + D) - 2 * x;
        const int y_ma = min(srows, 2 * y - D / 2 + D) - 2 * y;
#pragma unroll
        for(int yi = -2; yi <= 2; ++yi)
#pragma unroll
            for(int xi = -2; xi <= 2; ++xi)
            {
                const bool in_win = yi >= y_mi && yi < y_ma && xi >= x_mi && xi < x_ma;
                const int val = in_win ? (int)at(2 * y + yi, 2 * x + xi) : center;
                const int w = ((xi == 0) ? 6 : (xi == 1 || xi == -1) ? 4 : 1) * ((yi == 0) ? 6 : (yi == 1 || yi == -1) ? 4 : 1);
                const bool in = in_win && abs(val - center) < color_gate;
                isum += in ? val * w : 0;
                iwall += in ? w : 0;
            }
    }
    const float sum = (float)isum * 0.00390625f, wall = (float)iwall * 0.00390625f;
    return static_cast<int>(sum / wall);
}

__device__ __forceinline__ uint16_t pyr_down_u16_px(const uint16_t * __restrict__ src, int sp, int srows, int scols, int x, int y)
{
    return pyr_down_u16_at([&](int r, int c) { return (int)__ldg(src + (size_t)r * sp + c); }, srows, scols, x, y);
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

// {1,4,6,4,1}[k], 0 <= k <= 4 (cudafuncs.cu:453-457: the 5x5 kernel is the outer product of this row with itself)
__device__ __forceinline__ int gauss5_tap(int k) { return (0x14641 >> (4 * (k & 7))) & 15; }

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
    // border: the window [max(0, 2x-2), tx) x [max(0, 2y-2), ty) with tx / ty clamped to cols-1 / rows-1 (exclusive, :340-341)
    // and the kernel index counted back from that clamped end (:352) -- the same raster order as the reference loop, written
    // as 25 predicated taps so that a warp with one border pixel does not fall into a data-dependent loop
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    float sum = 0;
    int count = 0;
#pragma unroll
    for(int j = 0; j < 5; j++)
#pragma unroll
        for(int i = 0; i < 5; i++)
        {
            const int cy = 2 * y - 2 + j, cx = 2 * x - 2 + i;
            const bool in = cy >= 0 && cy < ty && cx >= 0 && cx < tx;
            const float s = in ? at(cy, cx) : qnan();
            if(!isnan(s))
            {
                const int w = gauss5_tap(ty - cy - 1) * gauss5_tap(tx - cx - 1);
                sum = __fmaf_rn(s, (float)w, sum); // the reference binary contracts `sum += s * k` into one FMA
                count += w;
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
    // border: 25 predicated taps, kernel index counted back from the clamped window end (see pyr_down_gauss_f32_px); the
    // float sums of the reference are exact small integers, so they are formed as integers
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int isum = 0, count = 0;
#pragma unroll
    for(int j = 0; j < 5; j++)
#pragma unroll
        for(int i = 0; i < 5; i++)
        {
            const int cy = 2 * y - 2 + j, cx = 2 * x - 2 + i;
            const bool in = cy >= 0 && cy < ty && cx >= 0 && cx < tx;
            const int s = in ? (int)at(cy, cx) : 0;
            if(s > 0)
            {
                const int w = gauss5_tap(ty - cy - 1) * gauss5_tap(tx - cx - 1);
                isum += s * w;
                count += w;
            }
        }
    const uint8_t r = ((float)isum / (float)count);
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

// applyKernel taps (cudafuncs.cu:615-621): gx = {a, 0, -a; b, 0, -b; a, 0, -a}, gy = {a, b, a; 0, 0, 0; -a, -b, -a} with
// a = 0.52201f, b = 0.79451f, k = row * 3 + column.  Evaluated from k (a sign times one of the two literals: exact, the same
// floats as the reference's table) instead of indexing a local array, which would live on the stack of every caller.
__device__ __forceinline__ float sobel_x_tap(int k)
{
    const int row = k / 3, col = k - row * 3;
    const float mag = (row == 1) ? 0.79451f : 0.52201f;
    return (col == 0) ? mag : ((col == 2) ? -mag : 0.f);
}
__device__ __forceinline__ float sobel_y_tap(int k)
{
    const int row = k / 3, col = k - row * 3;
    const float mag = (col == 1) ? 0.79451f : 0.52201f;
    return (row == 0) ? mag : ((row == 2) ? -mag : 0.f);
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
