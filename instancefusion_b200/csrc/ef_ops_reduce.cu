// ef_ops_reduce.cu -- Tier-2 association + reduction operators: icpStep, computeRgbResidual, rgbStep,
// so3Step (elasticfusionpublic/Core/src/Cuda/reduce.cu) as ONE launch each:
//   per-pixel math (ef_pixel.cuh) -> per-thread fp32 accumulation -> transpose-reduce warp butterfly ->
//   block partial row -> last-block-done final sum in block-index order (ef_reduce.cuh).
// The reference needs two launches (step kernel + reduceSum<<<1,1024>>>) and its sum order depends on
// the (threads, blocks) it is given; here the order is fixed by the image size alone.
//
// Memory behaviour (B200): current-frame maps are read with 128-bit loads, four consecutive pixels per
// thread, fully coalesced; the model maps are gathered through the read-only path (neighbouring
// pixels project to neighbouring model pixels, so the gather stays sector-coalesced).  Grid = enough
// 256-thread blocks for one pass, capped at a multiple of the SM count.
#include "ef_kernels.h"
#include "ef_pixel.cuh"
#include "ef_reduce.cuh"

namespace ef
{

namespace
{

constexpr int kBlock = 256;

inline Mat33 to_mat(const float * m)
{
    Mat33 r;
    r.r0 = make_float3(m[0], m[1], m[2]);
    r.r1 = make_float3(m[3], m[4], m[5]);
    r.r2 = make_float3(m[6], m[7], m[8]);
    return r;
}

__device__ __forceinline__ Mat33 mat_of(const float * m)
{
    Mat33 r;
    r.r0 = make_float3(m[0], m[1], m[2]);
    r.r1 = make_float3(m[3], m[4], m[5]);
    r.r2 = make_float3(m[6], m[7], m[8]);
    return r;
}

inline int num_sms()
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

inline int grid_for(int work_items)
{
    int blocks = (work_items + kBlock - 1) / kBlock;
    const int cap = num_sms() * 8; // 8 x 256 threads = 2048 resident threads per SM
    if(blocks > cap) blocks = cap;
    if(blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
    if(blocks < 1) blocks = 1;
    return blocks;
}

// ------------------------------------------------------------------------------------------------
// icpStep  (reduce.cu:257-490)
// ------------------------------------------------------------------------------------------------
// Without control flow in the pixel body: the PX pixels of a thread interleave, all their gathers are in flight
// together, rejected pixels gather pixel 0 and contribute rows of exact zeros (same bits as the branching form).
// ~220 instructions per 48 bytes: at 1280x720 the instruction issue (7 M warp instructions / 592 schedulers) costs
// about as much as the 44 MB from HBM, so the grid is sized to the real occupancy (no partial second wave) and the
// kernel runs grid-stride.
template<int PX>
__global__ void __launch_bounds__(kBlock, 2) k_icp_step(const IcpParams P0, const Map3 vc, const Map3 nc, const Map3 vp, const Map3 np,
                                                        int groups_per_row, int total_groups, float * __restrict__ partials,
                                                        unsigned * ticket, float * __restrict__ out, const IterParams * __restrict__ it)
{
    __shared__ float smem[32 * 32];
    // graph replay (EF_OPT_USE_GRAPH): the launch is a fixed graph node, the pose of the iteration comes from memory
    IcpParams P = P0;
    if(it)
    {
        P.Rcurr = mat_of(it->Rcurr);
        P.tcurr = make_float3(it->tcurr[0], it->tcurr[1], it->tcurr[2]);
        P.Rprev_inv = mat_of(it->Rprev_inv);
        P.tprev = make_float3(it->tprev[0], it->tprev[1], it->tprev[2]);
    }
    float acc[32];
#pragma unroll
    for(int i = 0; i < 32; i++) acc[i] = 0.f;
    const size_t plane = (size_t)vp.rows * vp.pitch; // the four maps share rows and pitch (launch_icp_step)

    for(int g = blockIdx.x * blockDim.x + threadIdx.x; g < total_groups; g += gridDim.x * blockDim.x)
    {
        const int y = g / groups_per_row;
        const int x0 = (g - y * groups_per_row) * PX;
        float vx[PX], vy[PX], vz[PX], nx[PX], ny[PX], nz[PX];
        if constexpr(PX == 4)
        {
            const float4 a = *reinterpret_cast<const float4 *>(vc.row(0, y) + x0);
            const float4 b = *reinterpret_cast<const float4 *>(vc.row(1, y) + x0);
            const float4 c = *reinterpret_cast<const float4 *>(vc.row(2, y) + x0);
            const float4 d = *reinterpret_cast<const float4 *>(nc.row(0, y) + x0);
            const float4 e = *reinterpret_cast<const float4 *>(nc.row(1, y) + x0);
            const float4 f = *reinterpret_cast<const float4 *>(nc.row(2, y) + x0);
            vx[0] = a.x; vx[1] = a.y; vx[2] = a.z; vx[3] = a.w;
            vy[0] = b.x; vy[1] = b.y; vy[2] = b.z; vy[3] = b.w;
            vz[0] = c.x; vz[1] = c.y; vz[2] = c.z; vz[3] = c.w;
            nx[0] = d.x; nx[1] = d.y; nx[2] = d.z; nx[3] = d.w;
            ny[0] = e.x; ny[1] = e.y; ny[2] = e.z; ny[3] = e.w;
            nz[0] = f.x; nz[1] = f.y; nz[2] = f.z; nz[3] = f.w;
        }
        else
        {
            vx[0] = vc.row(0, y)[x0]; vy[0] = vc.row(1, y)[x0]; vz[0] = vc.row(2, y)[x0];
            nx[0] = nc.row(0, y)[x0]; ny[0] = nc.row(1, y)[x0]; nz[0] = nc.row(2, y)[x0];
        }
        float3 vg[PX], mv[PX], mn[PX];
        bool in1[PX];
        size_t q[PX];
#pragma unroll
        for(int k = 0; k < PX; k++)
        {
            int ux, uy;
            in1[k] = icp_project(P, make_float3(vx[k], vy[k], vz[k]), vg[k], ux, uy);
            q[k] = in1[k] ? (size_t)uy * vp.pitch + ux : 0;
        }
#pragma unroll
        for(int k = 0; k < PX; k++)
        {
            mv[k].x = __ldg(vp.p + q[k]); mv[k].y = __ldg(vp.p + plane + q[k]); mv[k].z = __ldg(vp.p + 2 * plane + q[k]);
            mn[k].x = __ldg(np.p + q[k]); mn[k].y = __ldg(np.p + plane + q[k]); mn[k].z = __ldg(np.p + 2 * plane + q[k]);
        }
#pragma unroll
        for(int k = 0; k < PX; k++)
        {
            float row[7];
            const bool ok = icp_finish_select(P, vg[k], make_float3(nx[k], ny[k], nz[k]), mv[k], mn[k], row) && in1[k];
#pragma unroll
            for(int i = 0; i < 7; i++) row[i] = ok ? row[i] : 0.f;
            int kk = 0;
#pragma unroll
            for(int i = 0; i < 6; i++)
            {
#pragma unroll
                for(int j = i; j < 7; j++) acc[kk++] += row[i] * row[j];
            }
            acc[27] += row[6] * row[6];
            acc[28] += ok ? 1.0f : 0.f;
        }
    }

    const float lane_value = warp_transpose_reduce32(acc);
    const float block_total = block_reduce_slots<32>(lane_value, smem);
    grid_reduce_last_block<32>(block_total, partials, ticket, out, smem);
}

// blocks for `work_items` threads of `kernel`, at most one full wave of resident blocks (grid-stride kernels)
template<class K>
inline int grid_one_wave(K kernel, int work_items)
{
    int occ = 0;
    if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, 0) != cudaSuccess || occ < 1) occ = 1;
    int blocks = (work_items + kBlock - 1) / kBlock;
    const int cap = num_sms() * occ;
    if(blocks > cap) blocks = cap;
    if(blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
    if(blocks < 1) blocks = 1;
    return blocks;
}

// ------------------------------------------------------------------------------------------------
// computeRgbResidual  (reduce.cu:739-936): writes a 16-byte DataTerm per pixel (types.cuh:75-81) and
// reduces int2 {count, sum (int)(diff^2)} -- integer adds, so __reduce_add_sync + any order is exact.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_rgb_residual(const RgbResParams P0, const int16_t * __restrict__ dIdx,
                                                         const int16_t * __restrict__ dIdy, int d_pitch,
                                                         const float * __restrict__ last_depth, const float * __restrict__ next_depth,
                                                         int depth_pitch, const uint8_t * __restrict__ last_image,
                                                         const uint8_t * __restrict__ next_image, int img_pitch, int4 * __restrict__ corres,
                                                         int * __restrict__ partials, unsigned * ticket, int * __restrict__ out,
                                                         const IterParams * __restrict__ it)
{
    __shared__ int s_cnt[32], s_sig[32];
    __shared__ bool is_last;
    RgbResParams P = P0;
    if(it)
    {
        P.krkinv = mat_of(it->krkinv);
        P.kt = make_float3(it->kt[0], it->kt[1], it->kt[2]);
    }
    const int N = P.rows * P.cols;
    int cnt = 0, sig = 0;

    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int y = k / P.cols;
        const int x = k - y * P.cols;
        const int valx = dIdx[(size_t)y * d_pitch + x];
        const int valy = dIdy[(size_t)y * d_pitch + x];
        const float d1 = next_depth[(size_t)y * depth_pitch + x];
        int u0 = 0, v0 = 0;
        float diff = 0.f, d0;
        const bool ok = rgb_residual_px(P, x, y, valx, valy, d1, next_image, img_pitch, last_image, last_depth, depth_pitch, u0, v0, diff, d0);
        int4 rec;
        rec.x = (u0 & 0xffff) | (v0 << 16);  // short2 zero
        rec.y = (x & 0xffff) | (y << 16);    // short2 one
        rec.z = __float_as_int(diff);
        rec.w = ok ? 1 : 0;                  // bool valid (+3 pad bytes)
        corres[k] = rec;                     // written for every pixel (:839)
        if(ok)
        {
            cnt += 1;
            sig += (int)(diff * diff);       // :830 float -> int truncation
        }
    }

    cnt = __reduce_add_sync(kFullMask, cnt);
    sig = __reduce_add_sync(kFullMask, sig);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if(lane == 0) { s_cnt[warp] = cnt; s_sig[warp] = sig; }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        int c = 0, s = 0;
        for(unsigned w = 0; w < (blockDim.x >> 5); w++) { c += s_cnt[w]; s += s_sig[w]; }
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
        for(unsigned w = 0; w < (blockDim.x >> 5); w++) { c += s_cnt[w]; s += s_sig[w]; }
        out[0] = c;
        out[1] = s;
        *ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// rgbStep  (reduce.cu:494-678): consumes DataTerm records + the float3 point cloud
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_rgb_step(const RgbStepParams P0, const int4 * __restrict__ corres, const float * __restrict__ cloud,
                                                     int cloud_pitch /*floats*/, const int16_t * __restrict__ dIdx,
                                                     const int16_t * __restrict__ dIdy, int d_pitch, int N, float * __restrict__ partials,
                                                     unsigned * ticket, float * __restrict__ out, const IterParams * __restrict__ it,
                                                     const int * __restrict__ residual)
{
    __shared__ float smem[32 * 32];
    RgbStepParams P = P0;
    if(it)
    {
        // graph replay: the robust-weight scale from the {count, sum} computeRgbResidual left in device memory, formed
        // like the host does (RGBDOdometry.cpp:461, precedence quirk kept; :472)
        const int rgbSize = __ldcg(residual), sigma = __ldcg(residual + 1);
        float sigmaVal = (float)sqrt((double)((__fdiv_rn((float)sigma, (float)rgbSize) == 0) ? 1 : rgbSize));
        if(it->rgb_only) sigmaVal = -1;
        P.sigma = sigmaVal;
    }
    float acc[32];
#pragma unroll
    for(int i = 0; i < 32; i++) acc[i] = 0.f;

    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int4 rec = __ldg(corres + k);
        if(rec.w & 0xff)
        {
            const int u0 = (short)(rec.x & 0xffff), v0 = (short)(rec.x >> 16);
            const int x1 = (short)(rec.y & 0xffff), y1 = (short)(rec.y >> 16);
            const float diff = __int_as_float(rec.z);
            const float * c = cloud + (size_t)v0 * cloud_pitch + 3 * u0;
            float row[7];
            rgb_row(P, diff, __ldg(c), __ldg(c + 1), __ldg(c + 2), __ldg(dIdx + (size_t)y1 * d_pitch + x1),
                    __ldg(dIdy + (size_t)y1 * d_pitch + x1), row);
            accumulate_se3(acc, row);
        }
    }
    const float lane_value = warp_transpose_reduce32(acc);
    const float block_total = block_reduce_slots<32>(lane_value, smem);
    grid_reduce_last_block<32>(block_total, partials, ticket, out, smem);
}

// ------------------------------------------------------------------------------------------------
// so3Step  (reduce.cu:938-1141)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_so3_step(const So3Params P, const uint8_t * __restrict__ last_image,
                                                     const uint8_t * __restrict__ next_image, int pitch, float * __restrict__ partials,
                                                     unsigned * ticket, float * __restrict__ out)
{
    __shared__ float smem[32 * 16];
    float acc[16];
#pragma unroll
    for(int i = 0; i < 16; i++) acc[i] = 0.f;
    const int N = P.rows * P.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int y = k / P.cols;
        const int x = k - y * P.cols;
        float row[4];
        if(so3_row(P, x, y, last_image, next_image, pitch, row)) accumulate_so3(acc, row);
    }
    const float lane_value = warp_transpose_reduce16(acc);
    const float block_total = block_reduce_slots<16>(lane_value, smem);
    grid_reduce_last_block<16>(block_total, partials, ticket, out, smem);
}

} // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_icp_step(const IcpArgs & a, void * scratch, cudaStream_t s, const IterParams * it)
{
    IcpParams P;
    P.Rcurr = to_mat(a.Rcurr);
    P.tcurr = make_float3(a.tcurr[0], a.tcurr[1], a.tcurr[2]);
    P.Rprev_inv = to_mat(a.Rprev_inv);
    P.tprev = make_float3(a.tprev[0], a.tprev[1], a.tprev[2]);
    P.intr = Intr{a.fx, a.fy, a.cx, a.cy};
    P.dist_thresh = a.dist_thresh;
    P.angle_thresh = a.angle_thresh;
    P.rows = a.rows;
    P.cols = a.cols;
    const size_t pitch_b = a.pitch ? a.pitch : (size_t)a.cols * 4;
    const int pitch = (int)(pitch_b / 4);
    Map3 vc{a.vmap_curr, pitch, a.rows}, nc{a.nmap_curr, pitch, a.rows}, vp{a.vmap_g_prev, pitch, a.rows}, np{a.nmap_g_prev, pitch, a.rows};

    char * sc = static_cast<char *>(scratch);
    unsigned * ticket = reinterpret_cast<unsigned *>(sc + kScratchTicketOff);
    float * out = reinterpret_cast<float *>(sc + kScratchResultOff);
    float * partials = reinterpret_cast<float *>(sc + kScratchPartialOff);

    const bool vec = (a.cols % 4 == 0) && (pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.vmap_curr) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.nmap_curr) & 15) == 0);
    if(vec)
    {
        const int gpr = a.cols / 4, total = gpr * a.rows;
        k_icp_step<4><<<grid_one_wave(k_icp_step<4>, total), kBlock, 0, s>>>(P, vc, nc, vp, np, gpr, total, partials, ticket, out, it);
    }
    else
    {
        const int total = a.cols * a.rows;
        k_icp_step<1><<<grid_one_wave(k_icp_step<1>, total), kBlock, 0, s>>>(P, vc, nc, vp, np, a.cols, total, partials, ticket, out, it);
    }
    return cudaGetLastError();
}

cudaError_t launch_rgb_residual(const RgbResArgs & a, void * scratch, cudaStream_t s, const IterParams * it)
{
    RgbResParams P;
    P.krkinv = to_mat(a.krkinv);
    P.kt = make_float3(a.kt[0], a.kt[1], a.kt[2]);
    P.min_scale = a.min_scale;
    P.max_depth_delta = a.max_depth_delta;
    P.rows = a.rows;
    P.cols = a.cols;
    char * sc = static_cast<char *>(scratch);
    unsigned * ticket = reinterpret_cast<unsigned *>(sc + kScratchTicketOff);
    int * out = reinterpret_cast<int *>(sc + kScratchResultOff);
    int * partials = reinterpret_cast<int *>(sc + kScratchPartialOff);
    const int d_pitch = (int)((a.d_pitch ? a.d_pitch : (size_t)a.cols * 2) / 2);
    const int depth_pitch = (int)((a.depth_pitch ? a.depth_pitch : (size_t)a.cols * 4) / 4);
    const int img_pitch = (int)(a.image_pitch ? a.image_pitch : (size_t)a.cols);
    k_rgb_residual<<<grid_for(a.rows * a.cols), kBlock, 0, s>>>(P, a.dIdx, a.dIdy, d_pitch, a.last_depth, a.next_depth, depth_pitch,
                                                               a.last_image, a.next_image, img_pitch, static_cast<int4 *>(a.corres), partials,
                                                               ticket, out, it);
    return cudaGetLastError();
}

cudaError_t launch_rgb_step(const RgbStepArgs & a, void * scratch, cudaStream_t s, const IterParams * it, const int * residual)
{
    RgbStepParams P;
    P.sigma = a.sigma;
    P.fx = a.fx;
    P.fy = a.fy;
    P.inv_fx = 1.0f / a.fx;
    P.inv_fy = 1.0f / a.fy;
    P.cx = P.cy = 0.f;
    P.sobel_scale = a.sobel_scale;
    char * sc = static_cast<char *>(scratch);
    unsigned * ticket = reinterpret_cast<unsigned *>(sc + kScratchTicketOff);
    float * out = reinterpret_cast<float *>(sc + kScratchResultOff);
    float * partials = reinterpret_cast<float *>(sc + kScratchPartialOff);
    const int cloud_pitch = (int)((a.cloud_pitch ? a.cloud_pitch : (size_t)a.cols * 12) / 4);
    const int d_pitch = (int)((a.d_pitch ? a.d_pitch : (size_t)a.cols * 2) / 2);
    const int N = a.rows * a.cols;
    k_rgb_step<<<grid_for(N), kBlock, 0, s>>>(P, static_cast<const int4 *>(a.corres), a.cloud, cloud_pitch, a.dIdx, a.dIdy, d_pitch, N, partials,
                                             ticket, out, it, residual);
    return cudaGetLastError();
}

cudaError_t launch_so3_step(const So3Args & a, void * scratch, cudaStream_t s)
{
    So3Params P;
    P.image_basis = to_mat(a.image_basis);
    P.kinv = to_mat(a.kinv);
    P.krlr = to_mat(a.krlr);
    P.rows = a.rows;
    P.cols = a.cols;
    char * sc = static_cast<char *>(scratch);
    unsigned * ticket = reinterpret_cast<unsigned *>(sc + kScratchTicketOff);
    float * out = reinterpret_cast<float *>(sc + kScratchResultOff);
    float * partials = reinterpret_cast<float *>(sc + kScratchPartialOff);
    const int pitch = (int)(a.image_pitch ? a.image_pitch : (size_t)a.cols);
    k_so3_step<<<grid_for(a.rows * a.cols), kBlock, 0, s>>>(P, a.last_image, a.next_image, pitch, partials, ticket, out);
    return cudaGetLastError();
}

} // namespace ef
