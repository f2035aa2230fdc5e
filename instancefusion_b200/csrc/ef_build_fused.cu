// ef_build_fused.cu -- fused pyramid builders of the tracker handle (EF_OPT_FUSED_BUILD, default on).
//
// The reference builds its pyramids with one tiny kernel per operator and level: 33 launches, 8 device
// synchronisations and 4 cudaMalloc/cudaFree pairs per tracked frame (RGBDOdometry.cpp:118-265).  At 640x480
// every one of those kernels is launch-latency bound, so the builders here fuse along the data flow instead:
//
//   k_build_maps     initICP(maps) / initICPModel: RGBA32F vertex+normal textures -> 3-level SoA pyramids
//                    (copyMaps + 2x resizeVMap + 2x resizeNMap + 3x tranformMaps + the z channel kept for
//                    initRGB*) in ONE launch: a thread owns a 2x2 block, level 2 is formed by warp shuffles.
//   k_depth_level    initICP(depth), one launch per level: vertex map + normal map of the level (normals from
//                    vertices recomputed in registers) + the next level's bilateral pyrDown.
//   k_rgbd_level0/1  initRGB*: intensity + float depth at level 0 and their Gaussian pyrDowns.
//   k_derivatives3   computeDerivativeImages for the three levels in one launch.
//
// Per-pixel arithmetic comes from ef_image_px.cuh / ef_math.cuh, the same functions the stand-alone operator
// kernels use, so fused and per-operator paths produce identical bits (tests/test_tracker_gpu.py).
#include "ef_image_px.cuh"
#include "ef_kernels.h"

namespace ef
{

namespace
{

__device__ __forceinline__ void store3(float * base, size_t plane, size_t idx, bool valid, const float3 & v)
{
    const float q = qnan();
    base[idx] = valid ? v.x : q;
    base[plane + idx] = valid ? v.y : q;
    base[2 * plane + idx] = valid ? v.z : q;
}

__device__ __forceinline__ void store3x2(float * base, size_t plane, size_t idx, bool ok0, const float3 & a, bool ok1, const float3 & b)
{
    const float q = qnan();
    *reinterpret_cast<float2 *>(base + idx) = make_float2(ok0 ? a.x : q, ok1 ? b.x : q);
    *reinterpret_cast<float2 *>(base + plane + idx) = make_float2(ok0 ? a.y : q, ok1 ? b.y : q);
    *reinterpret_cast<float2 *>(base + 2 * plane + idx) = make_float2(ok0 ? a.z : q, ok1 ? b.z : q);
}

// ------------------------------------------------------------------------------------------------
// copyMaps (cudafuncs.cu:270-310) + resizeVMap/resizeNMap x2 (:365-444) + tranformMaps x3 (:206-268)
// lane = xb << 2 | yp << 1 | xp : the four lanes {xp, yp} of one xb hold the 2x2 level-1 block of a level-2 pixel
// ------------------------------------------------------------------------------------------------
struct MapsOut
{
    float * v[3];
    float * n[3];
};

template<bool TRANSFORM>
__global__ void __launch_bounds__(256) k_build_maps(const float4 * __restrict__ vsrc, const float4 * __restrict__ nsrc, int rows, int cols,
                                                    MapsOut out, float * __restrict__ tmp_z, Mat33 R, float3 t)
{
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int xp = lane & 1, yp = (lane >> 1) & 1, xb = lane >> 2;
    const int cols1 = cols >> 1, rows1 = rows >> 1, cols2 = cols >> 2, rows2 = rows >> 2;
    // a warp covers 16 x 2 level-1 pixels; blockDim.x = 256 = 8 warps side by side in x
    const int x1 = (blockIdx.x * 8 + warp_in_block) * 16 + xb * 2 + xp;
    const int y1 = blockIdx.y * 2 + yp;
    const bool inside = x1 < cols1 && y1 < rows1;

    float3 v1 = make_float3(0.f, 0.f, 0.f), n1 = v1;
    bool v1_ok = false, n1_ok = false;
    if(inside)
    {
        const size_t plane0 = (size_t)rows * cols;
        float3 vs[4], ns[4];
        bool ok[4];
#pragma unroll
        for(int r = 0; r < 2; r++)
        {
            const size_t o = (size_t)(2 * y1 + r) * cols + 2 * x1;
            const float4 a = __ldg(vsrc + o), b = __ldg(vsrc + o + 1), c = __ldg(nsrc + o), d = __ldg(nsrc + o + 1);
            vs[2 * r] = make_float3(a.x, a.y, a.z);
            vs[2 * r + 1] = make_float3(b.x, b.y, b.z);
            ns[2 * r] = make_float3(c.x, c.y, c.z);
            ns[2 * r + 1] = make_float3(d.x, d.y, d.z);
            ok[2 * r] = !(a.z == 0);     // both maps key on the VERTEX z (:285, :301)
            ok[2 * r + 1] = !(b.z == 0);
            // the reference keeps the whole RGBA32F vertex texture in vmaps_tmp; only its z channel is read again (:212)
            *reinterpret_cast<float2 *>(tmp_z + o) = make_float2(a.z, b.z);
        }
        // level 0 (copyMaps), transformed on the way out; two pixels of a row per 8-byte store
#pragma unroll
        for(int r = 0; r < 2; r++)
        {
            const size_t idx = (size_t)(2 * y1 + r) * cols + 2 * x1;
            float3 vo[2], no[2];
            bool nk[2];
#pragma unroll
            for(int i = 0; i < 2; i++)
            {
                const int k = 2 * r + i;
                vo[i] = vs[k];
                no[i] = ns[k];
                if(TRANSFORM)
                {
                    vo[i] = R * vs[k] + t;
                    no[i] = R * ns[k];
                }
                nk[i] = ok[k] && !isnan(ns[k].x);
            }
            store3x2(out.v[0], plane0, idx, ok[2 * r], vo[0], ok[2 * r + 1], vo[1]);
            store3x2(out.n[0], plane0, idx, nk[0], no[0], nk[1], no[1]);
        }
        // level 1 (resizeMapKernel): any NaN x among the four -> invalid; average in the order x00 + x01 + x10 + x11
        v1_ok = ok[0] && ok[1] && ok[2] && ok[3];
        n1_ok = v1_ok && !isnan(ns[0].x) && !isnan(ns[1].x) && !isnan(ns[2].x) && !isnan(ns[3].x);
        // (the vertex validity also needs non-NaN x, which holds for valid texels unless the caller passed NaNs)
        v1_ok = v1_ok && !isnan(vs[0].x) && !isnan(vs[1].x) && !isnan(vs[2].x) && !isnan(vs[3].x);
        v1.x = (vs[0].x + vs[1].x + vs[2].x + vs[3].x) / 4;
        v1.y = (vs[0].y + vs[1].y + vs[2].y + vs[3].y) / 4;
        v1.z = (vs[0].z + vs[1].z + vs[2].z + vs[3].z) / 4;
        n1.x = (ns[0].x + ns[1].x + ns[2].x + ns[3].x) / 4;
        n1.y = (ns[0].y + ns[1].y + ns[2].y + ns[3].y) / 4;
        n1.z = (ns[0].z + ns[1].z + ns[2].z + ns[3].z) / 4;
        n1 = normalized3(n1);
        const size_t plane1 = (size_t)rows1 * cols1;
        const size_t idx1 = (size_t)y1 * cols1 + x1;
        float3 vo = v1, no = n1;
        if(TRANSFORM)
        {
            vo = R * v1 + t;
            no = R * n1;
        }
        store3(out.v[1], plane1, idx1, v1_ok, vo);
        store3(out.n[1], plane1, idx1, n1_ok && !isnan(n1.x), no);
    }

    // level 2 from the four level-1 results of lanes base..base+3 (x00 = xp0 yp0, x01 = xp1 yp0, x10 = xp0 yp1, x11)
    const int base = lane & ~3;
    float3 vq[4], nq[4];
    bool vok = true, nok = true;
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
        vq[k].x = __shfl_sync(0xffffffffu, v1.x, base + k);
        vq[k].y = __shfl_sync(0xffffffffu, v1.y, base + k);
        vq[k].z = __shfl_sync(0xffffffffu, v1.z, base + k);
        nq[k].x = __shfl_sync(0xffffffffu, n1.x, base + k);
        nq[k].y = __shfl_sync(0xffffffffu, n1.y, base + k);
        nq[k].z = __shfl_sync(0xffffffffu, n1.z, base + k);
        const int vo = __shfl_sync(0xffffffffu, (int)v1_ok, base + k);
        const int no = __shfl_sync(0xffffffffu, (int)n1_ok, base + k);
        vok = vok && vo;
        nok = nok && no && !isnan(nq[k].x);
    }
    const int x2 = x1 >> 1, y2 = y1 >> 1;
    if((lane & 3) == 0 && x2 < cols2 && y2 < rows2)
    {
        float3 v2, n2;
        v2.x = (vq[0].x + vq[1].x + vq[2].x + vq[3].x) / 4;
        v2.y = (vq[0].y + vq[1].y + vq[2].y + vq[3].y) / 4;
        v2.z = (vq[0].z + vq[1].z + vq[2].z + vq[3].z) / 4;
        n2.x = (nq[0].x + nq[1].x + nq[2].x + nq[3].x) / 4;
        n2.y = (nq[0].y + nq[1].y + nq[2].y + nq[3].y) / 4;
        n2.z = (nq[0].z + nq[1].z + nq[2].z + nq[3].z) / 4;
        n2 = normalized3(n2);
        const size_t plane2 = (size_t)rows2 * cols2;
        const size_t idx2 = (size_t)y2 * cols2 + x2;
        float3 vo = v2, no = n2;
        if(TRANSFORM)
        {
            vo = R * v2 + t;
            no = R * n2;
        }
        store3(out.v[2], plane2, idx2, vok, vo);
        store3(out.n[2], plane2, idx2, nok && !isnan(n2.x), no);
    }
}

// ------------------------------------------------------------------------------------------------
// createVMap + createNMap of one level (+ pyrDown to the next).  A thread owns a 2x2 block of the level.
// ------------------------------------------------------------------------------------------------
template<bool HAS_NEXT, bool COPY_SRC>
__global__ void __launch_bounds__(256) k_depth_level(const uint16_t * __restrict__ depth, int dpitch /*elements*/, int rows, int cols, float fx_inv,
                                                     float fy_inv, float cx, float cy, float cutoff, float * __restrict__ vmap,
                                                     float * __restrict__ nmap, uint16_t * __restrict__ depth_copy,
                                                     uint16_t * __restrict__ next_depth)
{
    const int bx = blockIdx.x * blockDim.x + threadIdx.x; // block coordinates = next-level pixel
    const int by = blockIdx.y * blockDim.y + threadIdx.y;
    const int x0 = 2 * bx, y0 = 2 * by;
    if(x0 >= cols || y0 >= rows) return;
    const size_t plane = (size_t)rows * cols;

    // 3x3 depth neighbourhood (x0..x0+2, y0..y0+2), clamped reads for the unused out-of-image taps
    uint16_t d[3][3];
#pragma unroll
    for(int j = 0; j < 3; j++)
#pragma unroll
        for(int i = 0; i < 3; i++)
        {
            const int xx = min(x0 + i, cols - 1), yy = min(y0 + j, rows - 1);
            d[j][i] = __ldg(depth + (size_t)yy * dpitch + xx);
        }
    float3 v[3][3];
    bool ok[3][3];
#pragma unroll
    for(int j = 0; j < 3; j++)
#pragma unroll
        for(int i = 0; i < 3; i++) ok[j][i] = vertex_px(d[j][i], x0 + i, y0 + j, fx_inv, fy_inv, cx, cy, cutoff, v[j][i]);

#pragma unroll
    for(int j = 0; j < 2; j++)
#pragma unroll
        for(int i = 0; i < 2; i++)
        {
            const int x = x0 + i, y = y0 + j;
            if(x >= cols || y >= rows) continue;
            const size_t idx = (size_t)y * cols + x;
            if(COPY_SRC) depth_copy[idx] = d[j][i];
            // computeVmapKernel (:116-131): an invalid pixel only gets NaN in plane x
            if(ok[j][i])
            {
                vmap[idx] = v[j][i].x;
                vmap[plane + idx] = v[j][i].y;
                vmap[2 * plane + idx] = v[j][i].z;
            }
            else
                vmap[idx] = qnan();
            // computeNmapKernel (:159-187)
            if(x == cols - 1 || y == rows - 1 || !(ok[j][i] && ok[j][i + 1] && ok[j + 1][i]))
                nmap[idx] = qnan();
            else
            {
                const float3 n = normal_px(v[j][i], v[j][i + 1], v[j + 1][i]);
                nmap[idx] = n.x;
                nmap[plane + idx] = n.y;
                nmap[2 * plane + idx] = n.z;
            }
        }
    if(HAS_NEXT)
    {
        const int ncols = cols / 2, nrows = rows / 2;
        if(bx < ncols && by < nrows) next_depth[(size_t)by * ncols + bx] = pyr_down_u16_px(depth, dpitch, rows, cols, bx, by);
    }
}

// ------------------------------------------------------------------------------------------------
// populateRGBDData (RGBDOdometry.cpp:208-235), level 0 and level 1: a thread owns one level-1 pixel = a 2x2
// block of level 0.  The 5x5 windows of the pyrDowns re-derive their level-0 taps from the inputs.
// ------------------------------------------------------------------------------------------------
constexpr int kTileW = 2 * 32 + 3, kTileH = 2 * 8 + 3; // level-0 footprint of a 32x8 block of level-1 pixels

template<bool WITH_DEPTH>
__global__ void __launch_bounds__(256) k_rgbd_level0(const uint8_t * __restrict__ rgba, int rgba_pitch /*bytes*/, const float * __restrict__ tmp_z,
                                                     int z_stride, float cutoff, int rows, int cols, uint8_t * __restrict__ img0, float * __restrict__ depth0,
                                                     uint8_t * __restrict__ img1, float * __restrict__ depth1)
{
    __shared__ uint8_t s_img[kTileH][kTileW + 1];
    __shared__ float s_dep[WITH_DEPTH ? kTileH : 1][WITH_DEPTH ? kTileW : 1];
    const int cols1 = cols >> 1, rows1 = rows >> 1;
    const int ox = 2 * (int)(blockIdx.x * blockDim.x) - 2, oy = 2 * (int)(blockIdx.y * blockDim.y) - 2; // tile origin in level 0
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    // stage the level-0 intensity (and depth) of the whole footprint once: every texel is converted once per block
    for(int i = tid; i < kTileW * kTileH; i += 256)
    {
        const int ty = i / kTileW, tx = i - ty * kTileW;
        const int gx = ox + tx, gy = oy + ty;
        if(gx >= 0 && gy >= 0 && gx < cols && gy < rows)
        {
            s_img[ty][tx] = intensity_px(__ldg(reinterpret_cast<const uchar4 *>(rgba + (size_t)gy * rgba_pitch) + gx));
            if(WITH_DEPTH) s_dep[ty][tx] = depth_from_z(__ldg(tmp_z + ((size_t)gy * cols + gx) * z_stride), cutoff);
        }
    }
    __syncthreads();
    const int x1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y1 = blockIdx.y * blockDim.y + threadIdx.y;
    if(x1 >= cols1 || y1 >= rows1) return;
    auto inten = [&](int y, int x) { return s_img[y - oy][x - ox]; };
    auto dep = [&](int y, int x) { return s_dep[WITH_DEPTH ? y - oy : 0][WITH_DEPTH ? x - ox : 0]; };
#pragma unroll
    for(int j = 0; j < 2; j++)
    {
        const int y = 2 * y1 + j, x = 2 * x1;
        *reinterpret_cast<uchar2 *>(img0 + (size_t)y * cols + x) = make_uchar2(inten(y, x), inten(y, x + 1));
        if(WITH_DEPTH) *reinterpret_cast<float2 *>(depth0 + (size_t)y * cols + x) = make_float2(dep(y, x), dep(y, x + 1));
    }
    img1[(size_t)y1 * cols1 + x1] = pyr_down_gauss_u8_px(inten, rows, cols, x1, y1);
    if(WITH_DEPTH) depth1[(size_t)y1 * cols1 + x1] = pyr_down_gauss_f32_px(dep, rows, cols, x1, y1);
}

template<bool WITH_DEPTH>
__global__ void __launch_bounds__(256) k_rgbd_level1(const uint8_t * __restrict__ img1, const float * __restrict__ depth1, int rows1, int cols1,
                                                     uint8_t * __restrict__ img2, float * __restrict__ depth2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int cols2 = cols1 >> 1, rows2 = rows1 >> 1;
    if(x >= cols2 || y >= rows2) return;
    img2[(size_t)y * cols2 + x] = pyr_down_gauss_u8_px([&](int cy, int cx) { return __ldg(img1 + (size_t)cy * cols1 + cx); }, rows1, cols1, x, y);
    if(WITH_DEPTH)
        depth2[(size_t)y * cols2 + x] =
            pyr_down_gauss_f32_px([&](int cy, int cx) { return __ldg(depth1 + (size_t)cy * cols1 + cx); }, rows1, cols1, x, y);
}

// computeDerivativeImages for the three levels in one launch (cudafuncs.cu:583-639)
struct Deriv3
{
    const uint8_t * img[3];
    int16_t * dx[3];
    int16_t * dy[3];
    int rows[3], cols[3];
    int first_block[4]; // block ranges per level
};

__global__ void __launch_bounds__(256) k_derivatives3(Deriv3 D)
{
    int lvl = 0;
    if((int)blockIdx.x >= D.first_block[2]) lvl = 2;
    else if((int)blockIdx.x >= D.first_block[1]) lvl = 1;
    const int rows = D.rows[lvl], cols = D.cols[lvl];
    const int i = ((int)blockIdx.x - D.first_block[lvl]) * blockDim.x + threadIdx.x;
    if(i >= rows * cols) return;
    const int y = i / cols, x = i - y * cols;
    const uint8_t * img = D.img[lvl];
    short gx, gy;
    derivative_px([&](int j, int ii) { return __ldg(img + (size_t)j * cols + ii); }, rows, cols, x, y, gx, gy);
    D.dx[lvl][i] = gx;
    D.dy[lvl][i] = gy;
}

inline Mat33 to_mat(const float * m)
{
    Mat33 r;
    r.r0 = make_float3(m[0], m[1], m[2]);
    r.r1 = make_float3(m[3], m[4], m[5]);
    r.r2 = make_float3(m[6], m[7], m[8]);
    return r;
}

} // namespace

cudaError_t launch_build_maps(const float * v4, const float * n4, int rows, int cols, float * const vmaps[3], float * const nmaps[3], float * tmp_z,
                              const float * R, const float * t, cudaStream_t s)
{
    MapsOut out;
    for(int i = 0; i < 3; i++)
    {
        out.v[i] = vmaps[i];
        out.n[i] = nmaps[i];
    }
    const int cols1 = cols / 2, rows1 = rows / 2;
    const dim3 block(256);
    const dim3 grid((cols1 + 127) / 128, (rows1 + 1) / 2);
    if(R)
        k_build_maps<true><<<grid, block, 0, s>>>(reinterpret_cast<const float4 *>(v4), reinterpret_cast<const float4 *>(n4), rows, cols, out, tmp_z,
                                                 to_mat(R), make_float3(t[0], t[1], t[2]));
    else
    {
        Mat33 I;
        I.r0 = make_float3(1, 0, 0); I.r1 = make_float3(0, 1, 0); I.r2 = make_float3(0, 0, 1);
        k_build_maps<false><<<grid, block, 0, s>>>(reinterpret_cast<const float4 *>(v4), reinterpret_cast<const float4 *>(n4), rows, cols, out, tmp_z,
                                                  I, make_float3(0, 0, 0));
    }
    return cudaGetLastError();
}

cudaError_t launch_depth_level(const uint16_t * depth, size_t dpitch_bytes, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff,
                               float * vmap, float * nmap, uint16_t * depth_copy, uint16_t * next_depth, cudaStream_t s)
{
    const dim3 block(32, 8);
    const dim3 grid(((cols + 1) / 2 + 31) / 32, ((rows + 1) / 2 + 7) / 8);
    const int dp = (int)((dpitch_bytes ? dpitch_bytes : (size_t)cols * 2) / 2);
    const float fxi = 1.f / fx, fyi = 1.f / fy; // createVMap (:147)
    if(next_depth && depth_copy)
        k_depth_level<true, true><<<grid, block, 0, s>>>(depth, dp, rows, cols, fxi, fyi, cx, cy, cutoff, vmap, nmap, depth_copy, next_depth);
    else if(next_depth)
        k_depth_level<true, false><<<grid, block, 0, s>>>(depth, dp, rows, cols, fxi, fyi, cx, cy, cutoff, vmap, nmap, nullptr, next_depth);
    else if(depth_copy)
        k_depth_level<false, true><<<grid, block, 0, s>>>(depth, dp, rows, cols, fxi, fyi, cx, cy, cutoff, vmap, nmap, depth_copy, nullptr);
    else
        k_depth_level<false, false><<<grid, block, 0, s>>>(depth, dp, rows, cols, fxi, fyi, cx, cy, cutoff, vmap, nmap, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_rgbd_level0(const uint8_t * rgba, size_t pitch_bytes, const float * tmp_z, int z_stride, float cutoff, int rows, int cols, uint8_t * img0,
                               float * depth0, uint8_t * img1, float * depth1, cudaStream_t s)
{
    const dim3 block(32, 8);
    const dim3 grid((cols / 2 + 31) / 32, (rows / 2 + 7) / 8);
    const int p = (int)(pitch_bytes ? pitch_bytes : (size_t)cols * 4);
    if(depth0)
        k_rgbd_level0<true><<<grid, block, 0, s>>>(rgba, p, tmp_z, z_stride, cutoff, rows, cols, img0, depth0, img1, depth1);
    else
        k_rgbd_level0<false><<<grid, block, 0, s>>>(rgba, p, tmp_z, 1, cutoff, rows, cols, img0, nullptr, img1, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_rgbd_level1(const uint8_t * img1, const float * depth1, int rows1, int cols1, uint8_t * img2, float * depth2, cudaStream_t s)
{
    const dim3 block(32, 8);
    const dim3 grid((cols1 / 2 + 31) / 32, (rows1 / 2 + 7) / 8);
    if(depth1)
        k_rgbd_level1<true><<<grid, block, 0, s>>>(img1, depth1, rows1, cols1, img2, depth2);
    else
        k_rgbd_level1<false><<<grid, block, 0, s>>>(img1, nullptr, rows1, cols1, img2, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_derivatives3(const uint8_t * const img[3], int16_t * const dx[3], int16_t * const dy[3], const int rows[3], const int cols[3],
                                cudaStream_t s)
{
    Deriv3 D;
    int blocks = 0;
    for(int i = 0; i < 3; i++)
    {
        D.img[i] = img[i];
        D.dx[i] = dx[i];
        D.dy[i] = dy[i];
        D.rows[i] = rows[i];
        D.cols[i] = cols[i];
        D.first_block[i] = blocks;
        blocks += (rows[i] * cols[i] + 255) / 256;
    }
    D.first_block[3] = blocks;
    k_derivatives3<<<blocks, 256, 0, s>>>(D);
    return cudaGetLastError();
}

} // namespace ef
