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
#include <stdio.h>
#include <stdlib.h>
#include <vector>

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
__device__ __forceinline__ void build_maps_block(int block_x, int block_y, const float4 * __restrict__ vsrc, const float4 * __restrict__ nsrc, int rows,
                                                 int cols, const MapsOut & out, float * __restrict__ tmp_z, const Mat33 & R, const float3 & t)
{
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int xp = lane & 1, yp = (lane >> 1) & 1, xb = lane >> 2;
    const int cols1 = cols >> 1, rows1 = rows >> 1, cols2 = cols >> 2, rows2 = rows >> 2;
    // a warp covers 16 x 2 level-1 pixels; blockDim.x = 256 = 8 warps side by side in x
    const int x1 = (block_x * 8 + warp_in_block) * 16 + xb * 2 + xp;
    const int y1 = block_y * 2 + yp;
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

template<bool TRANSFORM>
__global__ void __launch_bounds__(256) k_build_maps(const float4 * __restrict__ vsrc, const float4 * __restrict__ nsrc, int rows, int cols,
                                                    MapsOut out, float * __restrict__ tmp_z, Mat33 R, float3 t)
{
    build_maps_block<TRANSFORM>(blockIdx.x, blockIdx.y, vsrc, nsrc, rows, cols, out, tmp_z, R, t);
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

// ------------------------------------------------------------------------------------------------
// k_build_frame: EVERY pyramid of a frame-to-model frame (ElasticFusion.cpp:343-368: initICPModel, initRGBModel,
// initICP(depth), initRGB) in ONE launch.  The level dependencies (level n+1 = pyrDown of level n) are resolved
// inside a CTA instead of between launches: a CTA owns a tile of kT2W x kT2H level-2 pixels, stages the level-0
// footprint of that tile plus the halo of the two 5x5 pyrDowns in shared memory, forms the level-1 footprint there,
// and writes the levels of its tile.  Halo pixels are evaluated by more than one CTA (1.5x of the taps, their loads
// are L2 hits) -- which costs less than the launch gaps and stream joins of chained kernels.  The kernel's duration
// is the instruction stream of its longest warp, so the work is cut into many short, independent jobs:
//   [intensity pyramid of the current image | of the model image | float depth pyramid | u16 depth levels 1, 2 with
//    their vertex / normal maps | vertex / normal map of level 0 | model map blocks (k_build_maps)]
// Per-pixel arithmetic: the same ef_image_px.cuh bodies as everywhere else.
// ------------------------------------------------------------------------------------------------
constexpr int kT2W = 16, kT2H = 8;                                 // level-2 pixels a tile owns
constexpr int kDep1W = 2 * kT2W + 5, kDep1H = 2 * kT2H + 5;        // u16 depth: level-1 footprint (normals need +1, pyrDown +-2)
constexpr int kDep0W = 2 * kDep1W + 3, kDep0H = 2 * kDep1H + 3;    // u16 depth: level-0 footprint
constexpr int kRgb1W = 2 * kT2W + 3, kRgb1H = 2 * kT2H + 3;        // Gaussian pyramids: level-1 footprint
constexpr int kRgb0W = 2 * kRgb1W + 3, kRgb0H = 2 * kRgb1H + 3;    // Gaussian pyramids: level-0 footprint
constexpr int kStageX = 8;                                         // the staged level-0 rows start kStageX pixels left of the tile ...
constexpr int kStageW = 4 * kT2W + 16;                             // ... and are this wide: groups of four pixels, aligned loads
static_assert(kStageX % 4 == 0 && kStageX >= 6 && kStageW % 4 == 0 && kStageW - kStageX >= 4 * kT2W + 7, "staged rows cover both footprints");
constexpr int kBuildJobs = 6;

struct FrameBuild
{
    int rows, cols;
    int tiles_x, tiles_y, maps_bx, maps_by;
    int first_block[kBuildJobs + 1];
    // map blocks
    const float4 * vsrc, * nsrc;
    MapsOut maps;
    float * tmp_z;
    Mat33 R;
    float3 t;
    // depth tiles
    const uint16_t * depth;
    int dpitch; // elements
    float cutoff;
    float fx_inv[3], fy_inv[3], cx[3], cy[3];
    float * vmap[3], * nmap[3];
    uint16_t * depth_out[3]; // dense copy of level 0, pyrDown levels 1 and 2
    // Gaussian pyramids: intensity of the current (0) and the model (1) image; ONE float depth pyramid written to both
    // the "next" and the "last" set (the reference derives both from the same vmaps_tmp, RGBDOdometry.cpp:212/239/245)
    const uint8_t * rgba[2];
    int rgba_pitch[2]; // bytes
    const float * z;
    int z_stride;
    float cutoff_rgb;
    uint8_t * img[2][3];
    float * dep_a[3], * dep_b[3];
    unsigned long long * timeline; // EF_BUILD_TIMELINE=1: {start ns, end ns, SM} of every block (diagnostic)
};

struct alignas(16) BuildSmem
{
    union alignas(16)
    {
        struct
        {
            uint16_t d0[kDep0H][kStageW];
            uint16_t d1[kDep1H][kDep1W + 1];
            uint16_t d2[kT2H + 1][kT2W + 2];
        } dep;
        struct
        {
            float z0[kRgb0H][kStageW];
            float z1[kRgb1H][kRgb1W];
        } zp;
        struct
        {
            uint8_t i0[kRgb0H][kStageW];
            uint8_t i1[kRgb1H][kRgb1W + 1];
        } im;
    };
};

struct TileGeom
{
    int rows, cols, rows1, cols1, rows2, cols2;
    int x2_0, y2_0, x1_0, y1_0, x0_0, y0_0; // first owned pixel per level
    int ox1, oy1, ox0, oy0;                 // origin of the level-1 / level-0 footprint
    __device__ __forceinline__ TileGeom(const FrameBuild & P, int tile)
    {
        const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
        rows = P.rows; cols = P.cols; rows1 = rows >> 1; cols1 = cols >> 1; rows2 = rows1 >> 1; cols2 = cols1 >> 1;
        x2_0 = tx * kT2W; y2_0 = ty * kT2H; x1_0 = 2 * x2_0; y1_0 = 2 * y2_0; x0_0 = 2 * x1_0; y0_0 = 2 * y1_0;
        ox1 = x1_0 - 2; oy1 = y1_0 - 2; ox0 = 2 * ox1 - 2; oy0 = 2 * oy1 - 2;
    }
    __device__ __forceinline__ bool own0(int gx, int gy) const { return gx >= x0_0 && gx < x0_0 + 4 * kT2W && gy >= y0_0 && gy < y0_0 + 4 * kT2H; }
    __device__ __forceinline__ bool own1(int gx, int gy) const { return gx >= x1_0 && gx < x1_0 + 2 * kT2W && gy >= y1_0 && gy < y1_0 + 2 * kT2H; }
};

// vertex + normal map of the pixels [x_lo, x_lo + w) x [y_lo, y_lo + h) of one level, the depth read through `at`
template<class F>
__device__ __forceinline__ void vn_from_depth(F at, int rows, int cols, int x_lo, int y_lo, int w, int h, float fxi, float fyi, float cx, float cy,
                                              float cutoff, float * __restrict__ vmap, float * __restrict__ nmap)
{
    const size_t plane = (size_t)rows * cols;
    for(int i = threadIdx.x; i < w * h; i += 256)
    {
        const int ly = i / w, lx = i - ly * w;
        const int x = x_lo + lx, y = y_lo + ly;
        if(x >= cols || y >= rows) continue;
        const size_t idx = (size_t)y * cols + x;
        const bool edge = x == cols - 1 || y == rows - 1;
        const uint16_t d00 = at(y, x), d01 = edge ? 0 : at(y, x + 1), d10 = edge ? 0 : at(y + 1, x);
        float3 v00, v01, v10;
        const bool ok00 = vertex_px(d00, x, y, fxi, fyi, cx, cy, cutoff, v00);
        const bool ok01 = vertex_px(d01, x + 1, y, fxi, fyi, cx, cy, cutoff, v01);
        const bool ok10 = vertex_px(d10, x, y + 1, fxi, fyi, cx, cy, cutoff, v10);
        // computeVmapKernel (:116-131): an invalid pixel only gets NaN in plane x
        if(ok00)
        {
            vmap[idx] = v00.x;
            vmap[plane + idx] = v00.y;
            vmap[2 * plane + idx] = v00.z;
        }
        else
            vmap[idx] = qnan();
        // computeNmapKernel (:159-187)
        if(ok00 && ok01 && ok10 && !edge)
        {
            const float3 n = normal_px(v00, v01, v10);
            nmap[idx] = n.x;
            nmap[plane + idx] = n.y;
            nmap[2 * plane + idx] = n.z;
        }
        else
            nmap[idx] = qnan();
    }
}

// initICP(depth), level 0: no pyrDown involved, three depth reads per pixel straight from the input
__device__ __forceinline__ void depth_level0_tile(const FrameBuild & P, int tile)
{
    const TileGeom G(P, tile);
    const uint16_t * depth = P.depth;
    const int dp = P.dpitch, cols = G.cols;
    uint16_t * copy = P.depth_out[0];
    vn_from_depth(
        [&](int r, int c) {
            const uint16_t d = __ldg(depth + (size_t)r * dp + c);
            return d;
        },
        G.rows, G.cols, G.x0_0, G.y0_0, 4 * kT2W, 4 * kT2H, P.fx_inv[0], P.fy_inv[0], P.cx[0], P.cy[0], P.cutoff, P.vmap[0], P.nmap[0]);
    // dense copy of the input (depth_tmp[0] of the reference)
    for(int i = threadIdx.x; i < 4 * kT2W * 4 * kT2H; i += 256)
    {
        const int ly = i / (4 * kT2W), lx = i - ly * (4 * kT2W);
        const int x = G.x0_0 + lx, y = G.y0_0 + ly;
        if(x < cols && y < G.rows) copy[(size_t)y * cols + x] = __ldg(depth + (size_t)y * dp + x);
    }
}

// initICP(depth), levels 1 and 2: bilateral pyrDown twice through shared memory, then the vertex / normal maps
__device__ __forceinline__ void depth_level12_tile(const FrameBuild & P, BuildSmem & S, int tile)
{
    const TileGeom G(P, tile);
    const int oy0 = G.oy0, ox1 = G.ox1, oy1 = G.oy1;
    // level 0 footprint, four pixels (one aligned 8-byte load) per item; the items of a thread are in flight together.
    // Rows are staged from column x0_0 - kStageX (a multiple of 4; the frame's width is one too).
    const int sx0 = G.x0_0 - kStageX;
    constexpr int kItems = (kStageW / 4) * kDep0H, kPer = (kItems + 255) / 256;
    {
        uint2 d[kPer];
#pragma unroll
        for(int k = 0; k < kPer; k++)
        {
            const int i = threadIdx.x + k * 256;
            const int ly = i / (kStageW / 4), lg = i - ly * (kStageW / 4);
            const int gx = sx0 + 4 * lg, gy = oy0 + ly;
            d[k] = make_uint2(0u, 0u);
            if(i < kItems && gx >= 0 && gy >= 0 && gx < G.cols && gy < G.rows)
                d[k] = __ldg(reinterpret_cast<const uint2 *>(P.depth + (size_t)gy * P.dpitch + gx));
        }
#pragma unroll
        for(int k = 0; k < kPer; k++)
        {
            const int i = threadIdx.x + k * 256;
            const int ly = i / (kStageW / 4), lg = i - ly * (kStageW / 4);
            if(i < kItems) *reinterpret_cast<uint2 *>(&S.dep.d0[ly][4 * lg]) = d[k];
        }
    }
    __syncthreads();
    auto at0 = [&](int r, int c) { return (int)S.dep.d0[r - oy0][c - sx0]; };
    for(int i = threadIdx.x; i < kDep1W * kDep1H; i += 256)
    {
        const int ly = i / kDep1W, lx = i - ly * kDep1W;
        const int gx = ox1 + lx, gy = oy1 + ly;
        uint16_t d = 0;
        if(gx >= 0 && gy >= 0 && gx < G.cols1 && gy < G.rows1)
        {
            d = pyr_down_u16_at(at0, G.rows, G.cols, gx, gy);
            if(G.own1(gx, gy)) P.depth_out[1][(size_t)gy * G.cols1 + gx] = d;
        }
        S.dep.d1[ly][lx] = d;
    }
    __syncthreads();
    auto at1 = [&](int r, int c) { return (int)S.dep.d1[r - oy1][c - ox1]; };
    for(int i = threadIdx.x; i < (kT2W + 1) * (kT2H + 1); i += 256)
    {
        const int ly = i / (kT2W + 1), lx = i - ly * (kT2W + 1);
        const int gx = G.x2_0 + lx, gy = G.y2_0 + ly;
        uint16_t d = 0;
        if(gx < G.cols2 && gy < G.rows2)
        {
            d = pyr_down_u16_at(at1, G.rows1, G.cols1, gx, gy);
            if(lx < kT2W && ly < kT2H) P.depth_out[2][(size_t)gy * G.cols2 + gx] = d;
        }
        S.dep.d2[ly][lx] = d;
    }
    vn_from_depth([&](int r, int c) { return S.dep.d1[r - oy1][c - ox1]; }, G.rows1, G.cols1, G.x1_0, G.y1_0, 2 * kT2W, 2 * kT2H, P.fx_inv[1],
                  P.fy_inv[1], P.cx[1], P.cy[1], P.cutoff, P.vmap[1], P.nmap[1]);
    __syncthreads();
    const int x2_0 = G.x2_0, y2_0 = G.y2_0;
    vn_from_depth([&](int r, int c) { return S.dep.d2[r - y2_0][c - x2_0]; }, G.rows2, G.cols2, x2_0, y2_0, kT2W, kT2H, P.fx_inv[2], P.fy_inv[2],
                  P.cx[2], P.cy[2], P.cutoff, P.vmap[2], P.nmap[2]);
}

// populateRGBDData (RGBDOdometry.cpp:208-235): one Gaussian pyramid of a tile.  DEPTH = false: intensity of image `job`;
// DEPTH = true: the float depth pyramid (z of the model vertices), written to both depth sets.
template<bool DEPTH>
__device__ __forceinline__ void gauss_pyramid_tile(const FrameBuild & P, BuildSmem & S, int job, int tile)
{
    const TileGeom G(P, tile);
    const int oy0 = G.oy0, ox1 = G.ox1, oy1 = G.oy1;
    const uint8_t * rgba = P.rgba[job];
    const int pitch = P.rgba_pitch[job];
    uint8_t * const * img = P.img[job];
    // level 0 of the footprint, converted once per CTA, four pixels per item (aligned 16-byte RGBA load, one packed
    // 4-byte store); owned pixels are stored.  The items of a thread are requested together (the inputs of a frame come
    // from HBM).  Rows are staged from column x0_0 - kStageX (a multiple of 4; the frame's width is one too).
    const int sx0 = G.x0_0 - kStageX;
    constexpr int kItems = (kStageW / 4) * kRgb0H, kPer = (kItems + 255) / 256;
    {
        uint4 px[DEPTH ? 1 : kPer];
        float zz[DEPTH ? kPer : 1][4];
#pragma unroll
        for(int k = 0; k < kPer; k++)
        {
            const int i = threadIdx.x + k * 256;
            const int ly = i / (kStageW / 4), lg = i - ly * (kStageW / 4);
            const int gx = sx0 + 4 * lg, gy = oy0 + ly;
            const bool in = i < kItems && gx >= 0 && gy >= 0 && gx < G.cols && gy < G.rows;
            if constexpr(DEPTH)
            {
#pragma unroll
                for(int c = 0; c < 4; c++) zz[k][c] = in ? __ldg(P.z + ((size_t)gy * G.cols + gx + c) * P.z_stride) : 0.f;
            }
            else
                px[k] = in ? __ldg(reinterpret_cast<const uint4 *>(rgba + (size_t)gy * pitch) + (gx >> 2)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for(int k = 0; k < kPer; k++)
        {
            const int i = threadIdx.x + k * 256;
            const int ly = i / (kStageW / 4), lg = i - ly * (kStageW / 4);
            const int gx = sx0 + 4 * lg, gy = oy0 + ly;
            if(i < kItems && gx >= 0 && gy >= 0 && gx < G.cols && gy < G.rows)
            {
                const bool own = G.own0(gx, gy); // tiles and groups are both aligned to 4: a group is owned as a whole
                if constexpr(DEPTH)
                {
                    const float4 d = make_float4(depth_from_z(zz[k][0], P.cutoff_rgb), depth_from_z(zz[k][1], P.cutoff_rgb),
                                                 depth_from_z(zz[k][2], P.cutoff_rgb), depth_from_z(zz[k][3], P.cutoff_rgb));
                    *reinterpret_cast<float4 *>(&S.zp.z0[ly][4 * lg]) = d;
                    if(own)
                    {
                        *reinterpret_cast<float4 *>(P.dep_a[0] + (size_t)gy * G.cols + gx) = d;
                        *reinterpret_cast<float4 *>(P.dep_b[0] + (size_t)gy * G.cols + gx) = d;
                    }
                }
                else
                {
                    const unsigned w[4] = {px[k].x, px[k].y, px[k].z, px[k].w};
                    unsigned packed = 0;
#pragma unroll
                    for(int c = 0; c < 4; c++)
                        packed |= (unsigned)intensity_px(make_uchar4(w[c] & 0xffu, (w[c] >> 8) & 0xffu, (w[c] >> 16) & 0xffu, w[c] >> 24)) << (8 * c);
                    *reinterpret_cast<unsigned *>(&S.im.i0[ly][4 * lg]) = packed;
                    if(own) *reinterpret_cast<unsigned *>(img[0] + (size_t)gy * G.cols + gx) = packed;
                }
            }
        }
    }
    __syncthreads();
    auto i0 = [&](int r, int c) { return S.im.i0[r - oy0][c - sx0]; };
    auto z0 = [&](int r, int c) { return S.zp.z0[r - oy0][c - sx0]; };
    for(int i = threadIdx.x; i < kRgb1W * kRgb1H; i += 256)
    {
        const int ly = i / kRgb1W, lx = i - ly * kRgb1W;
        const int gx = ox1 + lx, gy = oy1 + ly;
        if(gx >= 0 && gy >= 0 && gx < G.cols1 && gy < G.rows1)
        {
            const bool own = G.own1(gx, gy);
            if(DEPTH)
            {
                const float d = pyr_down_gauss_f32_px(z0, G.rows, G.cols, gx, gy);
                S.zp.z1[ly][lx] = d;
                if(own)
                {
                    P.dep_a[1][(size_t)gy * G.cols1 + gx] = d;
                    P.dep_b[1][(size_t)gy * G.cols1 + gx] = d;
                }
            }
            else
            {
                const uint8_t v = pyr_down_gauss_u8_px(i0, G.rows, G.cols, gx, gy);
                S.im.i1[ly][lx] = v;
                if(own) img[1][(size_t)gy * G.cols1 + gx] = v;
            }
        }
    }
    __syncthreads();
    auto i1 = [&](int r, int c) { return S.im.i1[r - oy1][c - ox1]; };
    auto z1 = [&](int r, int c) { return S.zp.z1[r - oy1][c - ox1]; };
    if(threadIdx.x < kT2W * kT2H)
    {
        const int ly = threadIdx.x / kT2W, lx = threadIdx.x - ly * kT2W;
        const int gx = G.x2_0 + lx, gy = G.y2_0 + ly;
        if(gx < G.cols2 && gy < G.rows2)
        {
            if(DEPTH)
            {
                const float d = pyr_down_gauss_f32_px(z1, G.rows1, G.cols1, gx, gy);
                P.dep_a[2][(size_t)gy * G.cols2 + gx] = d;
                P.dep_b[2][(size_t)gy * G.cols2 + gx] = d;
            }
            else
                img[2][(size_t)gy * G.cols2 + gx] = pyr_down_gauss_u8_px(i1, G.rows1, G.cols1, gx, gy);
        }
    }
}

__global__ void __launch_bounds__(256, 5) k_build_frame(const __grid_constant__ FrameBuild P)
{
    __shared__ BuildSmem S;
    const int b = blockIdx.x;
    unsigned long long t_begin = 0;
    if(P.timeline && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    if(b < P.first_block[1]) depth_level12_tile(P, S, b);
    else if(b < P.first_block[2]) gauss_pyramid_tile<true>(P, S, 0, b - P.first_block[1]);
    else if(b < P.first_block[3]) gauss_pyramid_tile<false>(P, S, 0, b - P.first_block[2]);
    else if(b < P.first_block[4]) gauss_pyramid_tile<false>(P, S, 1, b - P.first_block[3]);
    else if(b < P.first_block[5]) depth_level0_tile(P, b - P.first_block[4]);
    else
    {
        const int m = b - P.first_block[5];
        const int by = m / P.maps_bx, bx = m - by * P.maps_bx;
        build_maps_block<true>(bx, by, P.vsrc, P.nsrc, P.rows, P.cols, P.maps, P.tmp_z, P.R, P.t);
    }
    if(P.timeline)
    {
        __syncthreads();
        if(threadIdx.x == 0)
        {
            unsigned long long t_end;
            unsigned sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            P.timeline[3 * b] = t_begin;
            P.timeline[3 * b + 1] = t_end;
            P.timeline[3 * b + 2] = sm;
        }
    }
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

cudaError_t launch_build_frame(const FrameBuildArgs & a, cudaStream_t s)
{
    FrameBuild P;
    P.rows = a.rows;
    P.cols = a.cols;
    P.tiles_x = (a.cols + 4 * kT2W - 1) / (4 * kT2W);
    P.tiles_y = (a.rows + 4 * kT2H - 1) / (4 * kT2H);
    const int cols1 = a.cols / 2, rows1 = a.rows / 2;
    P.maps_bx = (cols1 + 127) / 128;
    P.maps_by = (rows1 + 1) / 2;
    const int tiles = P.tiles_x * P.tiles_y;
    for(int j = 0; j < kBuildJobs; j++) P.first_block[j] = j * tiles; // five tile jobs, then the map blocks
    P.first_block[kBuildJobs] = (kBuildJobs - 1) * tiles + P.maps_bx * P.maps_by;
    P.vsrc = reinterpret_cast<const float4 *>(a.v4);
    P.nsrc = reinterpret_cast<const float4 *>(a.n4);
    P.tmp_z = a.tmp_z;
    P.R = to_mat(a.R);
    P.t = make_float3(a.t[0], a.t[1], a.t[2]);
    P.depth = a.depth;
    P.dpitch = (int)((a.depth_pitch_bytes ? a.depth_pitch_bytes : (size_t)a.cols * 2) / 2);
    P.cutoff = a.depth_cutoff;
    P.rgba[0] = a.rgba;
    P.rgba[1] = a.model_rgba;
    P.rgba_pitch[0] = P.rgba_pitch[1] = a.cols * 4;
    P.z = a.v4 + 2; // the depth of BOTH RGB-D pyramids is the z of the model vertices (RGBDOdometry.cpp:212)
    P.z_stride = 4;
    P.cutoff_rgb = a.rgb_depth_cutoff;
    for(int i = 0; i < 3; i++)
    {
        P.maps.v[i] = a.vmap_g_prev[i];
        P.maps.n[i] = a.nmap_g_prev[i];
        P.fx_inv[i] = 1.f / a.fx[i]; // createVMap (:147)
        P.fy_inv[i] = 1.f / a.fy[i];
        P.cx[i] = a.cx[i];
        P.cy[i] = a.cy[i];
        P.vmap[i] = a.vmap_curr[i];
        P.nmap[i] = a.nmap_curr[i];
        P.depth_out[i] = a.depth_pyr[i];
        P.img[0][i] = a.next_image[i];
        P.img[1][i] = a.last_image[i];
        P.dep_a[i] = a.next_depth[i];
        P.dep_b[i] = a.last_depth[i];
    }
    P.timeline = nullptr;
    static const bool want_timeline = getenv("EF_BUILD_TIMELINE") && getenv("EF_BUILD_TIMELINE")[0] == '1';
    if(want_timeline)
    {
        // diagnostic: where every block ran and for how long; the table of the 20th launch goes to stderr
        static unsigned long long * d_tl = nullptr;
        static int calls = 0;
        const int nb = P.first_block[kBuildJobs];
        if(!d_tl) cudaMalloc((void **)&d_tl, (size_t)3 * 8192 * sizeof(unsigned long long));
        P.timeline = nb <= 8192 ? d_tl : nullptr;
        k_build_frame<<<nb, 256, 0, s>>>(P);
        if(P.timeline && ++calls == 20)
        {
            std::vector<unsigned long long> h((size_t)3 * nb);
            cudaStreamSynchronize(s);
            cudaMemcpy(h.data(), d_tl, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull, t1 = 0;
            for(int b = 0; b < nb; b++) { t0 = h[3 * b] < t0 ? h[3 * b] : t0; t1 = h[3 * b + 1] > t1 ? h[3 * b + 1] : t1; }
            fprintf(stderr, "[k_build_frame timeline] %d blocks, %.2f us from the first start to the last end\n", nb, (t1 - t0) * 1e-3);
            static const char * names[kBuildJobs] = {"u16 depth levels 1+2", "float depth pyramid", "intensity pyramid (current)", "intensity pyramid (model)", "vertex/normal level 0", "model map blocks"};
            for(int j = 0; j < kBuildJobs; j++)
            {
                double sum = 0, mx = 0, first = 1e30, last = 0;
                const int n = P.first_block[j + 1] - P.first_block[j];
                for(int b = P.first_block[j]; b < P.first_block[j + 1]; b++)
                {
                    const double d = (h[3 * b + 1] - h[3 * b]) * 1e-3;
                    sum += d; mx = d > mx ? d : mx;
                    first = (h[3 * b] - t0) * 1e-3 < first ? (h[3 * b] - t0) * 1e-3 : first;
                    last = (h[3 * b + 1] - t0) * 1e-3 > last ? (h[3 * b + 1] - t0) * 1e-3 : last;
                }
                fprintf(stderr, "  %-28s %4d blocks  mean %6.2f us  max %6.2f us  first start %6.2f  last end %6.2f  (block-time %7.1f us)\n", names[j], n, sum / n, mx, first, last, sum);
            }
        }
        return cudaGetLastError();
    }
    k_build_frame<<<P.first_block[kBuildJobs], 256, 0, s>>>(P);
    return cudaGetLastError();
}

} // namespace ef
