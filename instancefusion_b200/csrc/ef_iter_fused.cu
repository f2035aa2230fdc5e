// ef_iter_fused.cu -- EF_SOLVE_HOST, one Gauss-Newton iteration as TWO launches and no copy
// (elasticfusionpublic/Core/src/Utils/RGBDOdometry.cpp:436-540: computeRgbResidual, icpStep, rgbStep and their reduceSum
//  launches, 3 blocking downloads; Cuda/reduce.cu:257-936).
//
//   once per call   k_hm_gates        the iteration-invariant gates of RGBResidual::getProducts (reduce.cu:779-811: 16-pixel
//                                     border, gradient magnitude, depth validity, non-zero 4x4 window) for all three levels:
//                                     gate_depth(x, y) = nextDepth(x, y) where the pixel is a photometric candidate, NaN elsewhere
//   per iteration   k_hm_residual     warp + gather + accept of the candidates (reduce.cu:813-830) -> an 8-BYTE correspondence per
//                                     pixel {u | v << 12 | I_last << 24, lastDepth(u, v)} instead of the 16-byte DataTerm, and
//                                     {count, sum diff^2} by last-block-done
//                   k_hm_step         icpStep (reduce.cu:285-347) over the vertex / normal maps and rgbStep (reduce.cu:512-595)
//                                     over the correspondences -- the matched point is re-projected from its depth
//                                     (projectPointsKernel's expression, cudafuncs.cu:656-658) instead of being read from a 12-byte
//                                     point cloud --, the robust-weight scale formed from the residual kernel's sums
//                                     (RGBDOdometry.cpp:461), both 29-float sums reduced in one fixed order and stored, with the
//                                     residual sums and a sequence number, straight into mapped pinned memory: the host polls that
//                                     word instead of synchronising the stream.
// Four pixels per thread, 128-bit loads; same per-pixel expressions as the stand-alone operators (ef_pixel.cuh), so the
// correspondences and the integer sums are the ones k_rgb_residual produces and the float sums differ from the operators' only
// by the order of additions.
#include "ef_kernels.h"
#include "ef_pixel.cuh"
#include "ef_reduce.cuh"

namespace ef
{

namespace
{

constexpr int kBlock = 256;
constexpr unsigned kNoMatch = 0xffffffffu;

// accumulate_se3 without control flow: a rejected pixel adds a row of exact zeros and no inlier
__device__ __forceinline__ void accumulate_masked(float * acc, float * row, bool ok)
{
#pragma unroll
    for(int i = 0; i < 7; i++) row[i] = ok ? row[i] : 0.f;
    int k = 0;
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
#pragma unroll
        for(int j = i; j < 7; j++) acc[k++] += row[i] * row[j];
    }
    acc[27] += row[6] * row[6];
    acc[28] += ok ? 1.0f : 0.f;
}

struct Gates3
{
    const int16_t * dIdx[3], * dIdy[3];
    const float * next_depth[3];
    const uint8_t * next_image[3];
    float * gate_depth[3];
    float min_scale[3];
    int rows[3], cols[3];
    int first_block[4];
};

__global__ void __launch_bounds__(kBlock) k_hm_gates(const __grid_constant__ Gates3 G)
{
    const int lv = ((int)blockIdx.x >= G.first_block[1]) + ((int)blockIdx.x >= G.first_block[2]);
    const int rows = G.rows[lv], cols = G.cols[lv];
    const int k = ((int)blockIdx.x - G.first_block[lv]) * kBlock + (int)threadIdx.x;
    if(k >= rows * cols) return;
    const int y = k / cols, x = k - y * cols;
    RgbResParams P;
    P.min_scale = G.min_scale[lv];
    P.rows = rows;
    P.cols = cols;
    const float d1 = G.next_depth[lv][k];
    bool keep = rgb_gate(P, x, y, G.dIdx[lv][k], G.dIdy[lv][k], d1);
    if(keep)
    {
        // reduce.cu:787-793 (inside the 16-pixel border the clamps of the reference are no-ops)
        const uint8_t * img = G.next_image[lv];
#pragma unroll
        for(int u = -2; u < 2; u++)
        {
            const uint8_t * r = img + (size_t)(y + u) * cols + x;
#pragma unroll
            for(int v = -2; v < 2; v++) keep = keep && (__ldg(r + v) > 0);
        }
    }
    G.gate_depth[lv][k] = keep ? d1 : __int_as_float(0x7fc00000);
}

// reduce.cu:813-830 for four consecutive pixels of a row per thread
__global__ void __launch_bounds__(kBlock) k_hm_residual(const RgbResParams P, const float * __restrict__ gate_depth, const float * __restrict__ last_depth,
                                                        const uint8_t * __restrict__ last_image, const uint8_t * __restrict__ next_image,
                                                        unsigned * __restrict__ rec0, float * __restrict__ rec1, int * __restrict__ partials,
                                                        unsigned * ticket, int * __restrict__ out)
{
    __shared__ int s_cnt[kBlock / 32], s_sig[kBlock / 32];
    __shared__ bool is_last;
    const int gpr = P.cols / 4, total = gpr * P.rows;
    int cnt = 0, sig = 0;
    for(int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x)
    {
        const int y = g / gpr, x0 = (g - y * gpr) * 4;
        const size_t k0 = (size_t)y * P.cols + x0;
        const float4 dv = *reinterpret_cast<const float4 *>(gate_depth + k0);
        const unsigned inext = *reinterpret_cast<const unsigned *>(next_image + k0);
        const float d1[4] = {dv.x, dv.y, dv.z, dv.w};
        int u0[4], v0[4];
        float td1[4];
        bool ok[4];
        size_t q[4];
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
            ok[k] = rgb_warp(P, x0 + k, y, d1[k], u0[k], v0[k], td1[k]) && !isnan(d1[k]);
            q[k] = ok[k] ? (size_t)v0[k] * P.cols + u0[k] : 0;
        }
        float d0[4];
        unsigned l[4];
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
            d0[k] = __ldg(last_depth + q[k]);
            l[k] = __ldg(last_image + q[k]);
        }
        unsigned r0[4];
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
            const bool good = ok[k] && rgb_accept(P, td1[k], d0[k], (uint8_t)l[k]);
            const float diff = static_cast<float>((inext >> (8 * k)) & 0xffu) - static_cast<float>(l[k]); // :827
            r0[k] = good ? ((unsigned)u0[k] | ((unsigned)v0[k] << 12) | (l[k] << 24)) : kNoMatch;
            cnt += good ? 1 : 0;
            sig += good ? (int)(diff * diff) : 0; // :830
        }
        *reinterpret_cast<uint4 *>(rec0 + k0) = make_uint4(r0[0], r0[1], r0[2], r0[3]);
        *reinterpret_cast<float4 *>(rec1 + k0) = make_float4(d0[0], d0[1], d0[2], d0[3]);
    }
    // integer sums: exact in any order
    cnt = __reduce_add_sync(kFullMask, cnt);
    sig = __reduce_add_sync(kFullMask, sig);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if(lane == 0) { s_cnt[warp] = cnt; s_sig[warp] = sig; }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        int c = 0, s = 0;
        for(int w = 0; w < kBlock / 32; w++) { c += s_cnt[w]; s += s_sig[w]; }
        partials[blockIdx.x * 2 + 0] = c;
        partials[blockIdx.x * 2 + 1] = s;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if(!is_last) return;
    __threadfence();
    int c = 0, s = 0;
    for(unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
    {
        c += __ldcg(partials + b * 2 + 0);
        s += __ldcg(partials + b * 2 + 1);
    }
    c = __reduce_add_sync(kFullMask, c);
    s = __reduce_add_sync(kFullMask, s);
    __syncthreads();
    if(lane == 0) { s_cnt[warp] = c; s_sig[warp] = s; }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        c = 0; s = 0;
        for(int w = 0; w < kBlock / 32; w++) { c += s_cnt[w]; s += s_sig[w]; }
        out[0] = c;
        out[1] = s;
        *ticket = 0u;
    }
}

struct HmStep
{
    IcpParams icp;
    Map3 vc, nc, vp, np;
    RgbStepParams rgb;
    const unsigned * rec0;
    const float * rec1;
    const uint8_t * next_image;
    const int16_t * dIdx, * dIdy;
    const int * residual; // {count, sum diff^2} of this iteration's k_hm_residual
    int do_icp, do_rgb, rgb_only;
    int rows, cols;
    float * partials;     // gridDim.x rows of 64 floats
    unsigned * ticket;
    float * h_out;        // mapped pinned: [0, 29) ICP sums | [32, 61) RGB sums | [62] count [63] sum (int bits) | [64] sequence number
    unsigned seq;
};

__global__ void __launch_bounds__(kBlock, 2) k_hm_step(const __grid_constant__ HmStep A)
{
    __shared__ float smem[2][kBlock / 32][32];
    __shared__ bool is_last;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int gpr = A.cols / 4, total = gpr * A.rows;
    float vi = 0.f, vr = 0.f;

    if(A.do_icp)
    {
        // ---- icpStep: the body of k_icp_step<4> ----
        const IcpParams & P = A.icp;
        float acc[32];
#pragma unroll
        for(int i = 0; i < 32; i++) acc[i] = 0.f;
        const size_t plane = (size_t)A.vp.rows * A.vp.pitch;
        for(int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x)
        {
            const int y = g / gpr, x0 = (g - y * gpr) * 4;
            const float4 a = *reinterpret_cast<const float4 *>(A.vc.row(0, y) + x0);
            const float4 b = *reinterpret_cast<const float4 *>(A.vc.row(1, y) + x0);
            const float4 c = *reinterpret_cast<const float4 *>(A.vc.row(2, y) + x0);
            const float4 d = *reinterpret_cast<const float4 *>(A.nc.row(0, y) + x0);
            const float4 e = *reinterpret_cast<const float4 *>(A.nc.row(1, y) + x0);
            const float4 f = *reinterpret_cast<const float4 *>(A.nc.row(2, y) + x0);
            const float vx[4] = {a.x, a.y, a.z, a.w}, vy[4] = {b.x, b.y, b.z, b.w}, vz[4] = {c.x, c.y, c.z, c.w};
            const float nx[4] = {d.x, d.y, d.z, d.w}, ny[4] = {e.x, e.y, e.z, e.w}, nz[4] = {f.x, f.y, f.z, f.w};
            float3 vg[4], mv[4], mn[4];
            bool in1[4];
            size_t q[4];
#pragma unroll
            for(int k = 0; k < 4; k++)
            {
                int ux, uy;
                in1[k] = icp_project(P, make_float3(vx[k], vy[k], vz[k]), vg[k], ux, uy);
                q[k] = in1[k] ? (size_t)uy * A.vp.pitch + ux : 0;
            }
#pragma unroll
            for(int k = 0; k < 4; k++)
            {
                mv[k].x = __ldg(A.vp.p + q[k]); mv[k].y = __ldg(A.vp.p + plane + q[k]); mv[k].z = __ldg(A.vp.p + 2 * plane + q[k]);
                mn[k].x = __ldg(A.np.p + q[k]); mn[k].y = __ldg(A.np.p + plane + q[k]); mn[k].z = __ldg(A.np.p + 2 * plane + q[k]);
            }
#pragma unroll
            for(int k = 0; k < 4; k++)
            {
                float row[7];
                const bool ok = icp_finish_select(P, vg[k], make_float3(nx[k], ny[k], nz[k]), mv[k], mn[k], row) && in1[k];
                accumulate_masked(acc, row, ok);
            }
        }
        vi = warp_transpose_reduce32(acc);
    }

    if(A.do_rgb)
    {
        // ---- rgbStep over the 8-byte correspondences ----
        RgbStepParams P = A.rgb;
        {
            // RGBDOdometry.cpp:461 (precedence quirk kept), :472
            const int rgbSize = __ldcg(A.residual), sigma = __ldcg(A.residual + 1);
            float sigmaVal = (float)sqrt((double)((__fdiv_rn((float)sigma, (float)rgbSize) == 0) ? 1 : rgbSize));
            if(A.rgb_only) sigmaVal = -1;
            P.sigma = sigmaVal;
        }
        float acc[32];
#pragma unroll
        for(int i = 0; i < 32; i++) acc[i] = 0.f;
        for(int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x)
        {
            const int y = g / gpr, x0 = (g - y * gpr) * 4;
            const size_t k0 = (size_t)y * A.cols + x0;
            const uint4 r = __ldcg(reinterpret_cast<const uint4 *>(A.rec0 + k0));
            const unsigned rec[4] = {r.x, r.y, r.z, r.w};
            if((r.x & r.y & r.z & r.w) == kNoMatch) continue; // none of the four matched
            const float4 zv = __ldcg(reinterpret_cast<const float4 *>(A.rec1 + k0));
            const float Z[4] = {zv.x, zv.y, zv.z, zv.w};
            const unsigned inext = *reinterpret_cast<const unsigned *>(A.next_image + k0);
            const uint2 gxv = *reinterpret_cast<const uint2 *>(A.dIdx + k0), gyv = *reinterpret_cast<const uint2 *>(A.dIdy + k0);
            const short gx[4] = {(short)(gxv.x & 0xffffu), (short)(gxv.x >> 16), (short)(gxv.y & 0xffffu), (short)(gxv.y >> 16)};
            const short gy[4] = {(short)(gyv.x & 0xffffu), (short)(gyv.x >> 16), (short)(gyv.y & 0xffffu), (short)(gyv.y >> 16)};
#pragma unroll
            for(int k = 0; k < 4; k++)
            {
                const bool good = rec[k] != kNoMatch;
                const float diff = static_cast<float>((inext >> (8 * k)) & 0xffu) - static_cast<float>(rec[k] >> 24);
                const float3 cp = project_point((int)(rec[k] & 0xfffu), (int)((rec[k] >> 12) & 0xfffu), Z[k], P.inv_fx, P.inv_fy, P.cx, P.cy);
                float row[7];
                rgb_row(P, diff, cp.x, cp.y, cp.z, gx[k], gy[k], row);
                accumulate_masked(acc, row, good);
            }
        }
        vr = warp_transpose_reduce32(acc);
    }

    // ---- block totals (warps in index order), then the last block adds the block rows in index order ----
    smem[0][warp][lane] = vi;
    smem[1][warp][lane] = vr;
    __syncthreads();
    if(threadIdx.x < 64)
    {
        const int set = threadIdx.x >> 5;
        float tot = 0.f;
#pragma unroll
        for(int w = 0; w < kBlock / 32; w++) tot += smem[set][w][lane];
        A.partials[(size_t)blockIdx.x * 64 + threadIdx.x] = tot;
    }
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0) is_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if(!is_last) return;
    __threadfence();
    {
        // 4 parts x 64 slots; rows of a part requested 8 at a time, added in block order
        const unsigned slot = threadIdx.x & 63u, part = threadIdx.x >> 6;
        constexpr unsigned kParts = kBlock / 64, kTail = 8;
        float s = 0.f;
        for(unsigned b0 = part; b0 < gridDim.x; b0 += kTail * kParts)
        {
            float v[kTail];
#pragma unroll
            for(unsigned k = 0; k < kTail; k++)
            {
                const unsigned b = b0 + k * kParts;
                v[k] = b < gridDim.x ? __ldcg(A.partials + (size_t)b * 64 + slot) : 0.f;
            }
#pragma unroll
            for(unsigned k = 0; k < kTail; k++)
                if(b0 + k * kParts < gridDim.x) s += v[k];
        }
        float * red = &smem[0][0][0]; // 512 floats >= 4 x 64
        __syncthreads();
        red[part * 64 + slot] = s;
        __syncthreads();
        if(threadIdx.x < 64)
        {
            float tot = 0.f;
#pragma unroll
            for(unsigned p = 0; p < kParts; p++) tot += red[p * 64 + threadIdx.x];
            float o = tot;
            if(threadIdx.x == 62) o = __int_as_float(A.do_rgb ? __ldcg(A.residual) : 0);
            if(threadIdx.x == 63) o = __int_as_float(A.do_rgb ? __ldcg(A.residual + 1) : 0);
            A.h_out[threadIdx.x] = o;
        }
        __syncthreads();
        if(threadIdx.x == 0)
        {
            *A.ticket = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(A.h_out + 64) = A.seq;
            __threadfence_system();
        }
    }
}

int num_sms()
{
    static int n = 0;
    if(!n)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if(n <= 0) n = 148;
    }
    return n;
}

template<class K>
int one_wave(K kernel, int work_items, int max_blocks)
{
    static int occ = 0; // (one kernel per instantiation; asked once: the query costs as much as a launch)
    if(occ < 1 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, 0) != cudaSuccess || occ < 1)) occ = 1;
    int blocks = (work_items + kBlock - 1) / kBlock;
    const int cap = num_sms() * occ;
    if(blocks > cap) blocks = cap;
    if(blocks > max_blocks) blocks = max_blocks;
    if(blocks < 1) blocks = 1;
    return blocks;
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

cudaError_t launch_hm_gates(const HmGatesArgs & a, cudaStream_t s)
{
    Gates3 G;
    int blocks = 0;
    for(int i = 0; i < 3; i++)
    {
        G.dIdx[i] = a.dIdx[i]; G.dIdy[i] = a.dIdy[i];
        G.next_depth[i] = a.next_depth[i]; G.next_image[i] = a.next_image[i];
        G.gate_depth[i] = a.gate_depth[i];
        G.min_scale[i] = a.min_scale[i];
        G.rows[i] = a.rows[i]; G.cols[i] = a.cols[i];
        G.first_block[i] = blocks;
        blocks += (a.rows[i] * a.cols[i] + kBlock - 1) / kBlock;
    }
    G.first_block[3] = blocks;
    k_hm_gates<<<blocks, kBlock, 0, s>>>(G);
    return cudaGetLastError();
}

// scratch: one reduction scratch block (ef_kernels.h layout); result {count, sum} at scratch + kScratchResultOff
cudaError_t launch_hm_residual(const RgbResArgs & a, const float * gate_depth, unsigned * rec0, float * rec1, void * scratch, cudaStream_t s)
{
    RgbResParams P;
    P.krkinv = to_mat(a.krkinv);
    P.kt = make_float3(a.kt[0], a.kt[1], a.kt[2]);
    P.min_scale = a.min_scale;
    P.max_depth_delta = a.max_depth_delta;
    P.rows = a.rows;
    P.cols = a.cols;
    char * sc = static_cast<char *>(scratch);
    const int groups = a.rows * (a.cols / 4);
    int blocks = (groups + kBlock - 1) / kBlock;
    if(blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
    k_hm_residual<<<blocks, kBlock, 0, s>>>(P, gate_depth, a.last_depth, a.last_image, a.next_image, rec0, rec1,
                                           reinterpret_cast<int *>(sc + kScratchPartialOff), reinterpret_cast<unsigned *>(sc + kScratchTicketOff),
                                           reinterpret_cast<int *>(sc + kScratchResultOff));
    return cudaGetLastError();
}

// scratch: another reduction scratch block (partial rows of 64 floats: at most kMaxReduceBlocks / 2 blocks)
cudaError_t launch_hm_step(const HmStepArgs & a, void * scratch, cudaStream_t s)
{
    HmStep A;
    A.icp.Rcurr = to_mat(a.icp.Rcurr);
    A.icp.tcurr = make_float3(a.icp.tcurr[0], a.icp.tcurr[1], a.icp.tcurr[2]);
    A.icp.Rprev_inv = to_mat(a.icp.Rprev_inv);
    A.icp.tprev = make_float3(a.icp.tprev[0], a.icp.tprev[1], a.icp.tprev[2]);
    A.icp.intr = Intr{a.icp.fx, a.icp.fy, a.icp.cx, a.icp.cy};
    A.icp.dist_thresh = a.icp.dist_thresh;
    A.icp.angle_thresh = a.icp.angle_thresh;
    A.icp.rows = a.icp.rows;
    A.icp.cols = a.icp.cols;
    const int pitch = a.icp.cols;
    A.vc = Map3{a.icp.vmap_curr, pitch, a.icp.rows};
    A.nc = Map3{a.icp.nmap_curr, pitch, a.icp.rows};
    A.vp = Map3{a.icp.vmap_g_prev, pitch, a.icp.rows};
    A.np = Map3{a.icp.nmap_g_prev, pitch, a.icp.rows};
    A.rgb.sigma = 0.f;
    A.rgb.fx = a.icp.fx; A.rgb.fy = a.icp.fy;
    A.rgb.inv_fx = 1.0f / a.icp.fx; A.rgb.inv_fy = 1.0f / a.icp.fy;
    A.rgb.cx = a.icp.cx; A.rgb.cy = a.icp.cy;
    A.rgb.sobel_scale = a.sobel_scale;
    A.rec0 = a.rec0; A.rec1 = a.rec1;
    A.next_image = a.next_image;
    A.dIdx = a.dIdx; A.dIdy = a.dIdy;
    A.residual = a.residual;
    A.do_icp = a.do_icp; A.do_rgb = a.do_rgb; A.rgb_only = a.rgb_only;
    A.rows = a.icp.rows; A.cols = a.icp.cols;
    char * sc = static_cast<char *>(scratch);
    A.partials = reinterpret_cast<float *>(sc + kScratchPartialOff);
    A.ticket = reinterpret_cast<unsigned *>(sc + kScratchTicketOff);
    A.h_out = a.h_out;
    A.seq = a.seq;
    const int groups = a.icp.rows * (a.icp.cols / 4);
    k_hm_step<<<one_wave(k_hm_step, groups, kMaxReduceBlocks / 2), kBlock, 0, s>>>(A);
    return cudaGetLastError();
}

} // namespace ef
