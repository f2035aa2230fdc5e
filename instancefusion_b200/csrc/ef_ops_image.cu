// ef_ops_image.cu -- Tier-2 image / pyramid operators (one kernel per reference operator) on raw
// pitched images.  Citations: elasticfusionpublic/Core/src/Cuda/cudafuncs.cu.
//
// These are the per-operator drop-ins (and the building blocks of the non-fused tracker path);
// the fused pyramid builders used by the fast path live in ef_build_fused.cu.
#include "ef_kernels.h"
#include "ef_image_px.cuh"

namespace ef
{

namespace
{

template<class T> __device__ __forceinline__ const T * rowp(const T * base, size_t pitch_bytes, int y)
{
    return reinterpret_cast<const T *>(reinterpret_cast<const char *>(base) + (size_t)y * pitch_bytes);
}
template<class T> __device__ __forceinline__ T * rowp(T * base, size_t pitch_bytes, int y)
{
    return reinterpret_cast<T *>(reinterpret_cast<char *>(base) + (size_t)y * pitch_bytes);
}

inline dim3 grid2d(int cols, int rows, dim3 block) { return dim3((cols + block.x - 1) / block.x, (rows + block.y - 1) / block.y); }

// ---------------------------------------------------------------------------------------------
// pyrDownGaussKernel  cudafuncs.cu:57-94
// ---------------------------------------------------------------------------------------------
__global__ void k_pyr_down_u16(const uint16_t * __restrict__ src, size_t sp, int srows, int scols, uint16_t * __restrict__ dst,
                               size_t dp, int drows, int dcols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= dcols || y >= drows) return;
    rowp(dst, dp, y)[x] = pyr_down_u16_px(src, (int)(sp / 2), srows, scols, x, y);
}

// ---------------------------------------------------------------------------------------------
// computeVmapKernel  cudafuncs.cu:109-133
// ---------------------------------------------------------------------------------------------
__global__ void k_create_vmap(const uint16_t * __restrict__ depth, size_t dp, int rows, int cols, float fx_inv, float fy_inv, float cx,
                              float cy, float cutoff, float * __restrict__ vmap, size_t vp)
{
    const int u = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = threadIdx.y + blockIdx.y * blockDim.y;
    if(u < cols && v < rows)
    {
        float3 vtx;
        if(vertex_px(rowp(depth, dp, v)[u], u, v, fx_inv, fy_inv, cx, cy, cutoff, vtx))
        {
            rowp(vmap, vp, v)[u] = vtx.x;
            rowp(vmap, vp, v + rows)[u] = vtx.y;
            rowp(vmap, vp, v + rows * 2)[u] = vtx.z;
        }
        else
            rowp(vmap, vp, v)[u] = qnan(); // x plane only (:130)
    }
}

// ---------------------------------------------------------------------------------------------
// computeNmapKernel  cudafuncs.cu:151-188
// ---------------------------------------------------------------------------------------------
__global__ void k_create_nmap(int rows, int cols, const float * __restrict__ vmap, size_t vp, float * __restrict__ nmap, size_t np)
{
    const int u = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = threadIdx.y + blockIdx.y * blockDim.y;
    if(u >= cols || v >= rows) return;

    if(u == cols - 1 || v == rows - 1)
    {
        rowp(nmap, np, v)[u] = qnan();
        return;
    }

    float3 v00, v01, v10;
    v00.x = rowp(vmap, vp, v)[u];
    v01.x = rowp(vmap, vp, v)[u + 1];
    v10.x = rowp(vmap, vp, v + 1)[u];

    if(!isnan(v00.x) && !isnan(v01.x) && !isnan(v10.x))
    {
        v00.y = rowp(vmap, vp, v + rows)[u];
        v01.y = rowp(vmap, vp, v + rows)[u + 1];
        v10.y = rowp(vmap, vp, v + 1 + rows)[u];
        v00.z = rowp(vmap, vp, v + 2 * rows)[u];
        v01.z = rowp(vmap, vp, v + 2 * rows)[u + 1];
        v10.z = rowp(vmap, vp, v + 1 + 2 * rows)[u];

        const float3 r = normal_px(v00, v01, v10);
        rowp(nmap, np, v)[u] = r.x;
        rowp(nmap, np, v + rows)[u] = r.y;
        rowp(nmap, np, v + 2 * rows)[u] = r.z;
    }
    else
        rowp(nmap, np, v)[u] = qnan();
}

// ---------------------------------------------------------------------------------------------
// tranformMapsKernel  cudafuncs.cu:206-248 (src may alias dst: each thread reads then writes its own pixel)
// ---------------------------------------------------------------------------------------------
__global__ void k_transform_maps(int rows, int cols, const float * vsrc, const float * nsrc, size_t sp, Mat33 R, float3 t, float * vdst,
                                 float * ndst, size_t dp)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if(x >= cols || y >= rows) return;

    float3 vs, vd = make_float3(qnan(), qnan(), qnan());
    vs.x = rowp(vsrc, sp, y)[x];
    if(!isnan(vs.x))
    {
        vs.y = rowp(vsrc, sp, y + rows)[x];
        vs.z = rowp(vsrc, sp, y + 2 * rows)[x];
        vd = R * vs + t;
        rowp(vdst, dp, y + rows)[x] = vd.y;
        rowp(vdst, dp, y + 2 * rows)[x] = vd.z;
    }
    rowp(vdst, dp, y)[x] = vd.x;

    float3 ns, nd = make_float3(qnan(), qnan(), qnan());
    ns.x = rowp(nsrc, sp, y)[x];
    if(!isnan(ns.x))
    {
        ns.y = rowp(nsrc, sp, y + rows)[x];
        ns.z = rowp(nsrc, sp, y + 2 * rows)[x];
        nd = R * ns;
        rowp(ndst, dp, y + rows)[x] = nd.y;
        rowp(ndst, dp, y + 2 * rows)[x] = nd.z;
    }
    rowp(ndst, dp, y)[x] = nd.x;
}

// ---------------------------------------------------------------------------------------------
// copyMapsKernel  cudafuncs.cu:270-310   (one float4 load per map instead of three scalar loads)
// ---------------------------------------------------------------------------------------------
__global__ void k_copy_maps(int rows, int cols, const float4 * __restrict__ vsrc, const float4 * __restrict__ nsrc, float * __restrict__ vdst,
                            float * __restrict__ ndst, size_t dp)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if(x >= cols || y >= rows) return;

    const float4 vs = __ldg(vsrc + (size_t)y * cols + x);
    const float4 ns = __ldg(nsrc + (size_t)y * cols + x);
    const bool valid = !(vs.z == 0); // both maps key on the VERTEX z (:285, :301)
    const float q = qnan();
    rowp(vdst, dp, y)[x] = valid ? vs.x : q;
    rowp(vdst, dp, y + rows)[x] = valid ? vs.y : q;
    rowp(vdst, dp, y + 2 * rows)[x] = valid ? vs.z : q;
    rowp(ndst, dp, y)[x] = valid ? ns.x : q;
    rowp(ndst, dp, y + rows)[x] = valid ? ns.y : q;
    rowp(ndst, dp, y + 2 * rows)[x] = valid ? ns.z : q;
}

// ---------------------------------------------------------------------------------------------
// resizeMapKernel  cudafuncs.cu:365-416
// ---------------------------------------------------------------------------------------------
template<bool normalize>
__global__ void k_resize_map(int drows, int dcols, int srows, const float * __restrict__ in, size_t ip, float * __restrict__ out, size_t op)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if(x >= dcols || y >= drows) return;

    const int xs = x * 2, ys = y * 2;
    const float x00 = rowp(in, ip, ys + 0)[xs + 0];
    const float x01 = rowp(in, ip, ys + 0)[xs + 1];
    const float x10 = rowp(in, ip, ys + 1)[xs + 0];
    const float x11 = rowp(in, ip, ys + 1)[xs + 1];

    if(isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11))
    {
        rowp(out, op, y)[x] = qnan();
        return;
    }
    float3 n;
    n.x = (x00 + x01 + x10 + x11) / 4;
    const float y00 = rowp(in, ip, ys + srows + 0)[xs + 0];
    const float y01 = rowp(in, ip, ys + srows + 0)[xs + 1];
    const float y10 = rowp(in, ip, ys + srows + 1)[xs + 0];
    const float y11 = rowp(in, ip, ys + srows + 1)[xs + 1];
    n.y = (y00 + y01 + y10 + y11) / 4;
    const float z00 = rowp(in, ip, ys + 2 * srows + 0)[xs + 0];
    const float z01 = rowp(in, ip, ys + 2 * srows + 0)[xs + 1];
    const float z10 = rowp(in, ip, ys + 2 * srows + 1)[xs + 0];
    const float z11 = rowp(in, ip, ys + 2 * srows + 1)[xs + 1];
    n.z = (z00 + z01 + z10 + z11) / 4;
    if(normalize) n = normalized3(n);
    rowp(out, op, y)[x] = n.x;
    rowp(out, op, y + drows)[x] = n.y;
    rowp(out, op, y + 2 * drows)[x] = n.z;
}

// ---------------------------------------------------------------------------------------------
// verticesToDepthKernel  cudafuncs.cu:526-537
// ---------------------------------------------------------------------------------------------
__global__ void k_vertices_to_depth(const float * __restrict__ vsrc, int rows, int cols, float cutoff, float * __restrict__ dst, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= cols || y >= rows) return;
    const float z = __ldg(vsrc + ((size_t)y * cols + x) * 4 + 2);
    rowp(dst, dp, y)[x] = depth_from_z(z, cutoff);
}

// same predicate on a compact z image (the handle keeps only the z channel of vmaps_tmp)
__global__ void k_z_to_depth(const float * __restrict__ zsrc, int rows, int cols, float cutoff, float * __restrict__ dst, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= cols || y >= rows) return;
    const float z = __ldg(zsrc + (size_t)y * cols + x);
    rowp(dst, dp, y)[x] = depth_from_z(z, cutoff);
}

__global__ void k_extract_z(const float4 * __restrict__ vsrc, int n, float * __restrict__ z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) z[i] = __ldg(vsrc + i).z;
}

// ---------------------------------------------------------------------------------------------
// pyrDownKernelGaussF  cudafuncs.cu:332-363 ; binomial taps as constants instead of a malloc'd table
// ---------------------------------------------------------------------------------------------
__global__ void k_pyr_down_gauss_f32(const float * __restrict__ src, size_t sp, int srows, int scols, float * __restrict__ dst, size_t dp,
                                     int drows, int dcols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= dcols || y >= drows) return;
    rowp(dst, dp, y)[x] = pyr_down_gauss_f32_px([&](int cy, int cx) { return __ldg(rowp(src, sp, cy) + cx); }, srows, scols, x, y);
}

// ---------------------------------------------------------------------------------------------
// pyrDownKernelIntensityGauss  cudafuncs.cu:470-500
// ---------------------------------------------------------------------------------------------
__global__ void k_pyr_down_gauss_u8(const uint8_t * __restrict__ src, size_t sp, int srows, int scols, uint8_t * __restrict__ dst, size_t dp,
                                    int drows, int dcols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= dcols || y >= drows) return;
    rowp(dst, dp, y)[x] = pyr_down_gauss_u8_px([&](int cy, int cx) { return __ldg(rowp(src, sp, cy) + cx); }, srows, scols, x, y);
}

// ---------------------------------------------------------------------------------------------
// bgr2IntensityKernel  cudafuncs.cu:550-563 (uchar4 load from linear memory instead of tex2D)
// ---------------------------------------------------------------------------------------------
__global__ void k_bgr_to_intensity(const uint8_t * __restrict__ src, size_t sp, int rows, int cols, uint8_t * __restrict__ dst, size_t dp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= cols || y >= rows) return;
    rowp(dst, dp, y)[x] = intensity_px(__ldg(reinterpret_cast<const uchar4 *>(rowp(src, sp, y)) + x));
}

// ---------------------------------------------------------------------------------------------
// applyKernel  cudafuncs.cu:583-607 ; taps as literals (:615-621)
// ---------------------------------------------------------------------------------------------
__global__ void k_derivative_images(const uint8_t * __restrict__ src, size_t sp, int rows, int cols, int16_t * __restrict__ dx,
                                    int16_t * __restrict__ dy, size_t dp)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if(x >= cols || y >= rows) return;
    short gx, gy;
    derivative_px([&](int j, int i) { return __ldg(rowp(src, sp, j) + i); }, rows, cols, x, y, gx, gy);
    rowp(dx, dp, y)[x] = gx;
    rowp(dy, dp, y)[x] = gy;
}

// ---------------------------------------------------------------------------------------------
// projectPointsKernel  cudafuncs.cu:641-659
// ---------------------------------------------------------------------------------------------
__global__ void k_project_points(const float * __restrict__ depth, size_t dp, int rows, int cols, float * __restrict__ cloud, size_t cp,
                                 float invFx, float invFy, float cx, float cy)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if(x >= cols || y >= rows) return;
    const float z = rowp(depth, dp, y)[x];
    float * c = rowp(cloud, cp, y) + 3 * x;
    c[0] = (float)((x - cx) * z * invFx);
    c[1] = (float)((y - cy) * z * invFy);
    c[2] = z;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
#define EF_PITCH(p, dense) ((p) ? (p) : (size_t)(dense))

cudaError_t launch_pyr_down_u16(const uint16_t * src, size_t sp, int srows, int scols, uint16_t * dst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    const int drows = srows / 2, dcols = scols / 2;
    k_pyr_down_u16<<<grid2d(dcols, drows, block), block, 0, s>>>(src, EF_PITCH(sp, scols * 2), srows, scols, dst, EF_PITCH(dp, dcols * 2),
                                                                drows, dcols);
    return cudaGetLastError();
}

cudaError_t launch_create_vmap(const uint16_t * depth, size_t dp, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff,
                               float * vmap, size_t vp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_create_vmap<<<grid2d(cols, rows, block), block, 0, s>>>(depth, EF_PITCH(dp, cols * 2), rows, cols, 1.f / fx, 1.f / fy, cx, cy, cutoff,
                                                             vmap, EF_PITCH(vp, cols * 4));
    return cudaGetLastError();
}

cudaError_t launch_create_nmap(const float * vmap, size_t vp, int rows, int cols, float * nmap, size_t np, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_create_nmap<<<grid2d(cols, rows, block), block, 0, s>>>(rows, cols, vmap, EF_PITCH(vp, cols * 4), nmap, EF_PITCH(np, cols * 4));
    return cudaGetLastError();
}

cudaError_t launch_transform_maps(const float * vsrc, const float * nsrc, size_t sp, int rows, int cols, const float * R, const float * t,
                                  float * vdst, float * ndst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    Mat33 m;
    m.r0 = make_float3(R[0], R[1], R[2]);
    m.r1 = make_float3(R[3], R[4], R[5]);
    m.r2 = make_float3(R[6], R[7], R[8]);
    k_transform_maps<<<grid2d(cols, rows, block), block, 0, s>>>(rows, cols, vsrc, nsrc, EF_PITCH(sp, cols * 4), m,
                                                                make_float3(t[0], t[1], t[2]), vdst, ndst, EF_PITCH(dp, cols * 4));
    return cudaGetLastError();
}

cudaError_t launch_copy_maps(const float * v4, const float * n4, int rows, int cols, float * vdst, float * ndst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_copy_maps<<<grid2d(cols, rows, block), block, 0, s>>>(rows, cols, reinterpret_cast<const float4 *>(v4),
                                                           reinterpret_cast<const float4 *>(n4), vdst, ndst, EF_PITCH(dp, cols * 4));
    return cudaGetLastError();
}

cudaError_t launch_resize_map(const float * in, size_t ip, int srows, int scols, float * out, size_t op, int normalize, cudaStream_t s)
{
    const dim3 block(32, 8);
    const int drows = srows / 2, dcols = scols / 2;
    if(normalize)
        k_resize_map<true><<<grid2d(dcols, drows, block), block, 0, s>>>(drows, dcols, srows, in, EF_PITCH(ip, scols * 4), out,
                                                                        EF_PITCH(op, dcols * 4));
    else
        k_resize_map<false><<<grid2d(dcols, drows, block), block, 0, s>>>(drows, dcols, srows, in, EF_PITCH(ip, scols * 4), out,
                                                                         EF_PITCH(op, dcols * 4));
    return cudaGetLastError();
}

cudaError_t launch_vertices_to_depth(const float * v4, int rows, int cols, float cutoff, float * dst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_vertices_to_depth<<<grid2d(cols, rows, block), block, 0, s>>>(v4, rows, cols, cutoff, dst, EF_PITCH(dp, cols * 4));
    return cudaGetLastError();
}

cudaError_t launch_z_to_depth(const float * z, int rows, int cols, float cutoff, float * dst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_z_to_depth<<<grid2d(cols, rows, block), block, 0, s>>>(z, rows, cols, cutoff, dst, EF_PITCH(dp, cols * 4));
    return cudaGetLastError();
}

cudaError_t launch_extract_z(const float * v4, int n, float * z, cudaStream_t s)
{
    k_extract_z<<<(n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4 *>(v4), n, z);
    return cudaGetLastError();
}

cudaError_t launch_pyr_down_gauss_f32(const float * src, size_t sp, int srows, int scols, float * dst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    const int drows = srows / 2, dcols = scols / 2;
    k_pyr_down_gauss_f32<<<grid2d(dcols, drows, block), block, 0, s>>>(src, EF_PITCH(sp, scols * 4), srows, scols, dst,
                                                                      EF_PITCH(dp, dcols * 4), drows, dcols);
    return cudaGetLastError();
}

cudaError_t launch_pyr_down_gauss_u8(const uint8_t * src, size_t sp, int srows, int scols, uint8_t * dst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    const int drows = srows / 2, dcols = scols / 2;
    k_pyr_down_gauss_u8<<<grid2d(dcols, drows, block), block, 0, s>>>(src, EF_PITCH(sp, scols), srows, scols, dst, EF_PITCH(dp, dcols), drows,
                                                                     dcols);
    return cudaGetLastError();
}

cudaError_t launch_bgr_to_intensity(const uint8_t * rgba, size_t sp, int rows, int cols, uint8_t * dst, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_bgr_to_intensity<<<grid2d(cols, rows, block), block, 0, s>>>(rgba, EF_PITCH(sp, cols * 4), rows, cols, dst, EF_PITCH(dp, cols));
    return cudaGetLastError();
}

cudaError_t launch_derivative_images(const uint8_t * src, size_t sp, int rows, int cols, int16_t * dx, int16_t * dy, size_t dp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_derivative_images<<<grid2d(cols, rows, block), block, 0, s>>>(src, EF_PITCH(sp, cols), rows, cols, dx, dy, EF_PITCH(dp, cols * 2));
    return cudaGetLastError();
}

cudaError_t launch_project_points(const float * depth, size_t dp, int rows, int cols, float fx, float fy, float cx, float cy, float * cloud,
                                  size_t cp, cudaStream_t s)
{
    const dim3 block(32, 8);
    k_project_points<<<grid2d(cols, rows, block), block, 0, s>>>(depth, EF_PITCH(dp, cols * 4), rows, cols, cloud, EF_PITCH(cp, cols * 12),
                                                                1.0f / fx, 1.0f / fy, cx, cy);
    return cudaGetLastError();
}

} // namespace ef
