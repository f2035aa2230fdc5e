// ef_track_kernel.cu -- EF_SOLVE_DEVICE: the whole of RGBDOdometry::getIncrementalTransformation
// (elasticfusionpublic/Core/src/Utils/RGBDOdometry.cpp:292-591) as ONE persistent cooperative kernel.
//
// Why: at 640x480 a Gauss-Newton iteration touches ~19 MB that already sits in the 126 MB L2, i.e. a
// couple of microseconds of memory time, while the reference pays 3-4 launches, 3-4 device syncs and
// blocking D2H copies per iteration (57+ host round trips per frame).  Here one CTA per SM stays
// resident for the whole solve and an iteration costs a handful of L2 round trips:
//
//   phase A   every thread, for its 4-pixel groups: all coalesced loads first (current v/n maps as 128-bit
//             loads, gradients, depth, 4x4 validity window as 12 words), then both projections, then all
//             gathers, then the math: 29 ICP sums in registers, photometric correspondences -> SHARED memory.
//   barrier B (only with RGB: the weight needs the global count) ONE 64-bit atomic per CTA carries
//             arrivals | count | sum diff^2; whoever polls it gets all three in one load.
//   phase B   photometric rows from the records in shared memory -> 29 more sums.
//   reduce    transpose-reduce butterfly -> per-CTA 64-float partial row -> red.release arrival.
//   solve     CTA 0 sees the last arrival, adds the rows in CTA order (deterministic), ONE thread runs the
//             reference's host step in double (LDL^T, exp map, pose composition; ef_hostmath.h) and
//             publishes the next parameters in a 128-byte line whose four 32-byte sectors each carry the
//             epoch, so the other CTAs poll and load the parameters with the same single load.
//
// No DataTerm image, no point cloud, no reduceSum launch, no host involvement until the final pose is
// stored straight into pinned host memory.  Co-residency of the spinning CTAs is guaranteed by a
// cooperative launch with gridDim = number of SMs.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ef_hostmath.h"
#include "ef_kernels.h"
#include "ef_pixel.cuh"
#include "ef_reduce.cuh"
#include "ef_tracker.h"

namespace ef
{

namespace
{

constexpr int kThreads = 640;    // 20 warps = 5 per scheduler (96 registers each)
constexpr int kWarps = kThreads / 32;
constexpr int kMaxIters = 32;    // SE3 iterations per call (19 in the reference schedule)
constexpr int kDbgStamps = 10;
constexpr int kRowChunks = 20;   // a partial row = 20 x (3 floats + flag) = 60 floats >= 29 ICP + 29 RGB sums
constexpr int kRowFloats = kRowChunks * 3;
constexpr int kSo3Chunks = 4;    // SO3 rows carry 11 floats
constexpr int kLineChunks = 8;   // the parameter line = 8 x (3 floats + flag) = 24 floats
constexpr int kPayload = kLineChunks * 3;
constexpr int kParts = kThreads / 64; // final cross-CTA sum: 10 parts x 64 slots

struct LevelArgs
{
    const float * vc, * nc, * vp, * np;           // 3-plane maps, dense
    const float * last_depth, * next_depth;
    const uint8_t * last_image, * next_image;
    const int16_t * dIdx, * dIdy;
    int rows, cols;
    float fx, fy, cx, cy;                         // level intrinsics
    float inv_fx, inv_fy;                         // host 1.0f / f (cudafuncs.cu:671)
    float min_scale;
    int iterations;
    int px;                                       // pixels per thread unit at this level (4 or 1)
    double K_inv[9];                              // host double inverse of K (RGBDOdometry.cpp:428)
};

// Flagged 16-byte chunks (the NCCL "LL" idea): a chunk is written with ONE 128-bit store and read with ONE
// 128-bit load, so its three payload words and its flag are always observed together.  No fence, no separate
// "ready" counter: whoever polls a chunk gets the data with the same load that tells it the data is there.
// Flags are launch-unique epochs (launch_seq << 8 | n), so nothing has to be reset between launches.
struct alignas(128) TrackCtl
{
    uint4 line[kLineChunks];                      // parameters of the next phase, written by the solver thread
    unsigned long long bar_b[kMaxIters];          // per SE3 iteration: arrivals | count << 8 | sigma << 32
    double last_S[27];                            // combined normal equations of the last solve (for lastA / lastb)
};
// SE3 payload: Rcurr[9] tcurr[3] krkinv[9] kt[3];  SO3 payload: H[9] krlr[9] done

struct TrackOutput // pinned host memory, written by the solver thread
{
    float trans[3], rot[9];
    ef_track_stats st;
    int status;
};

struct TrackArgs
{
    LevelArgs lvl[kNumPyrs];
    const uint8_t * so3_last, * so3_next;         // level-2 lastNextImage / nextImage
    float so3_kinv[9];                            // (float) of the double inverse of K at level 2 (:323-325)
    float Rprev[9], tprev[3], Rprev_inv[9];
    float dist_thresh, angle_thresh, max_depth_delta, sobel_scale, icp_weight;
    int icp, rgb, rgb_only, so3;
    float prev_icp_error, prev_icp_count, prev_so3_error, prev_so3_count, prev_rgb_error, prev_rgb_count;
    unsigned epoch_base;                          // launch_seq << 8
    TrackCtl * ctl;
    uint4 * rows;                                 // (gridDim.x - 1) * kRowChunks flagged chunks, one row per worker CTA
    TrackOutput * out;
    long long * dbg;                              // optional clock64 stamps of CTA 0 (EF_TRACK_TIMING=1)
};

// ---- 128-bit relaxed (L2-coherent) accesses ----
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 * p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint4 * p, const uint4 & v)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// solver thread: eight self-validating chunks
__device__ __forceinline__ void publish_line(uint4 * line, const float * payload, unsigned epoch)
{
#pragma unroll
    for(int c = 0; c < kLineChunks; c++)
        st_relaxed_v4(line + c, make_uint4(__float_as_uint(payload[3 * c]), __float_as_uint(payload[3 * c + 1]), __float_as_uint(payload[3 * c + 2]), epoch));
}

// all threads call: lanes 0..7 of warp 0 spin on their chunk until it shows `epoch`; payload -> smem
__device__ __forceinline__ void wait_line(const uint4 * line, unsigned epoch, float * s_payload)
{
    if(threadIdx.x < 32)
    {
        const unsigned lane = threadIdx.x;
        uint4 v = make_uint4(0, 0, 0, epoch);
        do
        {
            if(lane < kLineChunks) v = ld_relaxed_v4(line + lane);
        } while(!__all_sync(kFullMask, v.w == epoch));
        if(lane < kLineChunks)
        {
            s_payload[3 * lane] = __uint_as_float(v.x);
            s_payload[3 * lane + 1] = __uint_as_float(v.y);
            s_payload[3 * lane + 2] = __uint_as_float(v.z);
        }
    }
    __syncthreads();
}

// worker CTA: publish this CTA's partial sums (already in shared memory) as flagged chunks
__device__ __forceinline__ void publish_row(uint4 * my_row, const float * s_row, int chunks, unsigned epoch)
{
    if((int)threadIdx.x < chunks)
    {
        const int c = threadIdx.x;
        st_relaxed_v4(my_row + c, make_uint4(__float_as_uint(s_row[3 * c]), __float_as_uint(s_row[3 * c + 1]), __float_as_uint(s_row[3 * c + 2]), epoch));
    }
}

// CTA 0: collect every worker's row (poll + load in one), then add the rows in worker order -> s_final[64]
__device__ __forceinline__ void gather_rows(const uint4 * rows, int workers, int chunks, unsigned epoch, float * s_rows /*[workers][kRowFloats]*/,
                                            float * s_red, float * s_final)
{
    const int total = workers * chunks;
    for(int i = threadIdx.x; i < total; i += kThreads)
    {
        const int w = i / chunks, c = i - w * chunks;
        const uint4 * p = rows + (size_t)w * kRowChunks + c;
        uint4 v;
        do
        {
            v = ld_relaxed_v4(p);
        } while(v.w != epoch);
        float * d = s_rows + w * kRowFloats + 3 * c;
        d[0] = __uint_as_float(v.x);
        d[1] = __uint_as_float(v.y);
        d[2] = __uint_as_float(v.z);
    }
    __syncthreads();
    const int nfl = chunks * 3;
    const int slot = threadIdx.x & 63, part = threadIdx.x >> 6;
    float s = 0.f;
    if(slot < nfl)
        for(int w = part; w < workers; w += kParts) s += s_rows[w * kRowFloats + slot];
    s_red[part * 64 + slot] = s;
    __syncthreads();
    if(threadIdx.x < 64)
    {
        float tot = 0.f;
#pragma unroll
        for(int p = 0; p < kParts; p++) tot += s_red[p * 64 + threadIdx.x];
        s_final[threadIdx.x] = tot;
    }
    __syncthreads();
}

__device__ __forceinline__ Mat33 mat_from(const float * m)
{
    Mat33 r;
    r.r0 = make_float3(m[0], m[1], m[2]);
    r.r1 = make_float3(m[3], m[4], m[5]);
    r.r2 = make_float3(m[6], m[7], m[8]);
    return r;
}

// ------------------------------------------------------------------------------------------------
// solver state kept by thread 0 of CTA 0 across iterations
// ------------------------------------------------------------------------------------------------
struct Solver
{
    double resultRt[16];
    float Rcurr[9], tcurr[3];
    float last_icp_error, last_icp_count, last_rgb_error, last_rgb_count, last_so3_error, last_so3_count;
    int so3_iterations, se3_iterations[3];
};

// RGBDOdometry.cpp:515-516, :541-583 -- runs in ONE thread (thread 0 of CTA 0).  Kept out of line so its
// double-precision register needs do not inflate the per-pixel phases of the kernel.  s_final (shared memory):
// ICP accumulator [0, 29) followed by the RGB accumulator [29, 58).
__device__ __noinline__ void solve_se3(Solver & S, const float * s_final, double * last_S, int icp, int rgb, float icp_weight,
                                       const float * Rprev, const float * tprev, int level)
{
    if(icp)
    {
        S.last_icp_error = sqrtf(s_final[27]) / s_final[28]; // :515-516
        S.last_icp_count = s_final[28];
    }
    double Sm[27];
    const double w = icp_weight;
    if(icp && rgb) // :547-553  A = A_rgb + w^2 A_icp ; b = b_rgb + w b_icp
    {
#pragma unroll
        for(int i = 0; i < 6; i++)
        {
#pragma unroll
            for(int j = i; j < 7; j++)
            {
                const int k = hm::acc_index(i, j);
                Sm[k] = (j == 6) ? ((double)s_final[29 + k] + w * (double)s_final[k]) : ((double)s_final[29 + k] + w * w * (double)s_final[k]);
            }
        }
    }
    else
    {
        const int o = icp ? 0 : 29;
#pragma unroll
        for(int k = 0; k < 27; k++) Sm[k] = (double)s_final[o + k];
    }
#pragma unroll
    for(int k = 0; k < 27; k++) last_S[k] = Sm[k];
    double result[6];
    hm::ldlt_solve_spd6_acc(Sm, result);
    S.se3_iterations[level]++;
    hm::update_se3(S.resultRt, result);                           // :573
    hm::compose_pose(S.resultRt, Rprev, tprev, S.Rcurr, S.tcurr); // :575-583
}

// :424-434 -- parameters of the next SE3 iteration (solver thread) -> payload[28]
__device__ __noinline__ void make_se3_payload(const Solver & S, float * payload, int rgb, float fx, float fy, float cx, float cy,
                                              const double * K_inv)
{
#pragma unroll
    for(int i = 0; i < 9; i++) payload[i] = S.Rcurr[i];
#pragma unroll
    for(int i = 0; i < 3; i++) payload[9 + i] = S.tcurr[i];
    if(rgb)
    {
        const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
        hm::rgb_warp_params(S.resultRt, K, K_inv, payload + 12, payload + 21);
    }
    else
    {
#pragma unroll
        for(int i = 12; i < 24; i++) payload[i] = 0.f;
    }
#pragma unroll
    for(int i = 24; i < kPayload; i++) payload[i] = 0.f;
}

struct So3State
{
    double resultR[9], lastResultR[9];
    float R_lr[9], lastError, lastCount;
};

// :348-380 -- digest one so3Step evaluation; returns done
__device__ __noinline__ int solve_so3(Solver & S, So3State & Z, const float * s_final, int it)
{
    int done = 0;
    float jtj[9], jtr[3], residual[2];
    hm::unpack_so3(s_final, jtj, jtr, residual);
    S.so3_iterations++;
    S.last_so3_error = sqrtf(residual[0]) / residual[1];                        // :348
    S.last_so3_count = residual[1];
    if(S.last_so3_error < Z.lastError && Z.lastCount == S.last_so3_count) done = 1; // :352
    else if(S.last_so3_error > Z.lastError + 0.001)                              // :356
    {
        S.last_so3_error = Z.lastError;
        S.last_so3_count = Z.lastCount;
#pragma unroll
        for(int i = 0; i < 9; i++) Z.resultR[i] = Z.lastResultR[i];
        done = 1;
    }
    if(!done)
    {
        Z.lastError = S.last_so3_error;
        Z.lastCount = S.last_so3_count;
#pragma unroll
        for(int i = 0; i < 9; i++) Z.lastResultR[i] = Z.resultR[i];
        float delta[3];
        hm::ldlt_solve<float, 3>(jtj, jtr, delta);                               // :368
        const double dd[3] = {delta[0], delta[1], delta[2]};
        double rotUpdate[9];
        hm::rodrigues(dd, rotUpdate);
        float ru[9];
#pragma unroll
        for(int i = 0; i < 9; i++) ru[i] = (float)rotUpdate[i];
        hm::mul33(ru, Z.R_lr, Z.R_lr);                                            // :372
#pragma unroll
        for(int i = 0; i < 9; i++) Z.resultR[i] = Z.R_lr[i];
        if(it == 10) done = 1; // ten evaluations made
    }
    return done;
}

// :318-329 -- homography K R K^-1 and K R for the next so3Step -> payload[28]
__device__ __noinline__ void make_so3_payload(const So3State & Z, float * payload, int done, float fx, float fy, float cx, float cy,
                                              const double * K_inv)
{
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    double KR[9], H[9];
    hm::mul33(K, Z.resultR, KR);
    hm::mul33(KR, K_inv, H);
#pragma unroll
    for(int i = 0; i < 9; i++)
    {
        payload[i] = (float)H[i];
        payload[9 + i] = (float)KR[i];
    }
    payload[18] = done ? 1.f : 0.f;
#pragma unroll
    for(int i = 19; i < kPayload; i++) payload[i] = 0.f;
}

// bytes k+2 .. k+5 of the 12-byte run (a0 a1 a2) all 0xFF ?  (4x4 validity window of pixel k of a group)
__device__ __forceinline__ bool window_ok(unsigned a0, unsigned a1, unsigned a2, int k)
{
    unsigned m;
    if(k == 0) m = __byte_perm(a0, a1, 0x5432);
    else if(k == 1) m = __byte_perm(a0, a1, 0x6543);
    else if(k == 2) m = a1;
    else m = __byte_perm(a1, a2, 0x4321);
    return m == 0xffffffffu;
}

// ------------------------------------------------------------------------------------------------
// per-unit pixel work.  A "unit" is PX consecutive pixels of one row handled by one thread (PX = 4 with
// 128-bit loads at the fine levels, PX = 1 at the coarse ones so that the few pixels spread over many threads).
// ------------------------------------------------------------------------------------------------
template<int PX>
__device__ __forceinline__ void rgb_assoc_unit(const LevelArgs & L, const RgbResParams & RP, int y, int x0, int4 * s_corr, int rec_base, int & cnt,
                                               int & sig)
{
    const int cols = L.cols;
    // the 16-pixel border of RGBResidual (:779-783); x0 is a multiple of PX so a 4-pixel unit is all in or all out
    const bool in_region = y >= 16 && y < L.rows - 16 && x0 >= 16 && x0 < cols - 16;
    short gxs[PX], gys[PX];
    float d1s[PX];
    unsigned ni4 = 0, m0 = 0, m1 = 0, m2 = 0;
    const int xa = x0 & ~3; // aligned start of the 12-byte validity run [xa-4, xa+8)
    if(in_region)
    {
        const size_t o = (size_t)y * cols + x0;
        if constexpr(PX == 4)
        {
            const short4 gx4 = *reinterpret_cast<const short4 *>(L.dIdx + o);
            const short4 gy4 = *reinterpret_cast<const short4 *>(L.dIdy + o);
            const float4 d14 = *reinterpret_cast<const float4 *>(L.next_depth + o);
            gxs[0] = gx4.x; gxs[1] = gx4.y; gxs[2] = gx4.z; gxs[3] = gx4.w;
            gys[0] = gy4.x; gys[1] = gy4.y; gys[2] = gy4.z; gys[3] = gy4.w;
            d1s[0] = d14.x; d1s[1] = d14.y; d1s[2] = d14.z; d1s[3] = d14.w;
        }
        else
        {
            gxs[0] = L.dIdx[o];
            gys[0] = L.dIdy[o];
            d1s[0] = L.next_depth[o];
        }
        // 4 rows x 12 bytes of the next image: non-zero masks, ANDed over the rows (:787-793)
        m0 = m1 = m2 = 0xffffffffu;
#pragma unroll
        for(int r = -2; r < 2; r++)
        {
            const unsigned * wp = reinterpret_cast<const unsigned *>(L.next_image + (size_t)(y + r) * cols + xa - 4);
            const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
            m0 &= __vcmpne4(w0, 0u);
            m1 &= __vcmpne4(w1, 0u);
            m2 &= __vcmpne4(w2, 0u);
            if(r == 0) ni4 = w1;
        }
    }
    int u0[PX], v0[PX];
    float td1[PX];
    bool ok[PX];
#pragma unroll
    for(int k = 0; k < PX; k++)
    {
        const int kk = (x0 + k) & 3;
        ok[k] = in_region && rgb_gate(RP, x0 + k, y, gxs[k], gys[k], d1s[k]) && window_ok(m0, m1, m2, kk);
        if(ok[k]) ok[k] = rgb_warp(RP, x0 + k, y, d1s[k], u0[k], v0[k], td1[k]);
    }
    float d0s[PX];
    unsigned ls[PX];
#pragma unroll
    for(int k = 0; k < PX; k++)
    {
        if(ok[k])
        {
            const size_t q = (size_t)v0[k] * cols + u0[k];
            d0s[k] = __ldg(L.last_depth + q);
            ls[k] = __ldg(L.last_image + q);
        }
    }
#pragma unroll
    for(int k = 0; k < PX; k++)
    {
        const int kk = (x0 + k) & 3;
        const bool good = ok[k] && rgb_accept(RP, td1[k], d0s[k], (uint8_t)ls[k]);
        int4 rec;
        rec.x = -1;
        rec.y = rec.z = rec.w = 0;
        if(good)
        {
            const float diff = static_cast<float>((ni4 >> (8 * kk)) & 0xffu) - static_cast<float>(ls[k]); // reduce.cu:827
            rec.x = (u0[k] & 0xffff) | (v0[k] << 16);
            rec.y = __float_as_int(diff);
            rec.z = __float_as_int(d0s[k]);
            rec.w = ((int)gxs[k] & 0xffff) | ((int)gys[k] << 16);
            cnt += 1;
            sig += (int)(diff * diff); // reduce.cu:830
        }
        s_corr[(rec_base + k) * kThreads + threadIdx.x] = rec;
    }
}

template<int PX>
__device__ __forceinline__ void icp_unit(const LevelArgs & L, const IcpParams & IP, int y, int x0, float * accI)
{
    const int cols = L.cols;
    const size_t plane = (size_t)L.rows * cols;
    const size_t o = (size_t)y * cols + x0;
    float vx[PX], vy[PX], vz[PX], nx[PX], ny[PX], nz[PX];
    if constexpr(PX == 4)
    {
        const float4 a = *reinterpret_cast<const float4 *>(L.vc + o);
        const float4 b = *reinterpret_cast<const float4 *>(L.vc + plane + o);
        const float4 c = *reinterpret_cast<const float4 *>(L.vc + 2 * plane + o);
        const float4 d = *reinterpret_cast<const float4 *>(L.nc + o);
        const float4 e = *reinterpret_cast<const float4 *>(L.nc + plane + o);
        const float4 f = *reinterpret_cast<const float4 *>(L.nc + 2 * plane + o);
        vx[0] = a.x; vx[1] = a.y; vx[2] = a.z; vx[3] = a.w;
        vy[0] = b.x; vy[1] = b.y; vy[2] = b.z; vy[3] = b.w;
        vz[0] = c.x; vz[1] = c.y; vz[2] = c.z; vz[3] = c.w;
        nx[0] = d.x; nx[1] = d.y; nx[2] = d.z; nx[3] = d.w;
        ny[0] = e.x; ny[1] = e.y; ny[2] = e.z; ny[3] = e.w;
        nz[0] = f.x; nz[1] = f.y; nz[2] = f.z; nz[3] = f.w;
    }
    else
    {
        vx[0] = L.vc[o]; vy[0] = L.vc[plane + o]; vz[0] = L.vc[2 * plane + o];
        nx[0] = L.nc[o]; ny[0] = L.nc[plane + o]; nz[0] = L.nc[2 * plane + o];
    }
    float3 vg[PX], vp[PX], np[PX];
    int ux[PX], uy[PX];
    bool in1[PX];
#pragma unroll
    for(int k = 0; k < PX; k++) in1[k] = icp_project(IP, make_float3(vx[k], vy[k], vz[k]), vg[k], ux[k], uy[k]);
#pragma unroll
    for(int k = 0; k < PX; k++)
    {
        if(in1[k])
        {
            const size_t q = (size_t)uy[k] * cols + ux[k];
            vp[k].x = __ldg(L.vp + q); vp[k].y = __ldg(L.vp + plane + q); vp[k].z = __ldg(L.vp + 2 * plane + q);
            np[k].x = __ldg(L.np + q); np[k].y = __ldg(L.np + plane + q); np[k].z = __ldg(L.np + 2 * plane + q);
        }
    }
#pragma unroll
    for(int k = 0; k < PX; k++)
    {
        float row[7];
        if(in1[k] && icp_finish(IP, vg[k], make_float3(nx[k], ny[k], nz[k]), vp[k], np[k], row)) accumulate_se3(accI, row);
    }
}

template<int PX>
__device__ __forceinline__ void rgb_rows_unit(const RgbStepParams & SP, const int4 * s_corr, int rec_base, float * accR)
{
#pragma unroll
    for(int k = 0; k < PX; k++)
    {
        const int4 rec = s_corr[(rec_base + k) * kThreads + threadIdx.x];
        if(rec.x != -1)
        {
            const int pu = rec.x & 0xffff, pv = rec.x >> 16;
            const float Z = __int_as_float(rec.z);
            const float3 cp = project_point(pu, pv, Z, SP.inv_fx, SP.inv_fy, SP.cx, SP.cy);
            float row[7];
            rgb_row(SP, __int_as_float(rec.y), cp.x, cp.y, cp.z, (short)(rec.w & 0xffff), (short)(rec.w >> 16), row);
            accumulate_se3(accR, row);
        }
    }
}

// Units are dealt to the worker CTAs in 32-unit chunks, round-robin, so that image regions with no work (the
// photometric border, depth holes) are spread evenly: chunk c -> worker c % W, handled by warp (c / W) % kWarps
// in pass (c / W) / kWarps.
struct UnitIter
{
    int units, W, worker, lane, warp;
    __device__ __forceinline__ int unit(int pass) const
    {
        const int local_chunk = pass * kWarps + warp;
        const int u = (local_chunk * W + worker) * 32 + lane;
        return u < units ? u : -1;
    }
    __device__ __forceinline__ int passes() const
    {
        const int chunks = (units + 31) / 32;
        const int per_worker = (chunks + W - 1) / W;
        return (per_worker + kWarps - 1) / kWarps;
    }
};

__global__ void __launch_bounds__(kThreads, 1) k_track(const TrackArgs A)
{
    extern __shared__ int4 s_dyn[];             // workers: correspondence records; CTA 0: gathered rows
    __shared__ float s_red[kWarps * 64];
    __shared__ float s_final[64];
    __shared__ float s_par[kPayload];
    __shared__ int s_cnt, s_sig;
    __shared__ int s_wcnt[kWarps], s_wsig[kWarps];

    TrackCtl * ctl = A.ctl;
    const unsigned grid = gridDim.x;
    const int W = (int)grid - 1;                // worker CTAs (blockIdx 1 .. grid-1); CTA 0 only gathers and solves
    const bool is_solver_cta = (blockIdx.x == 0);
    const bool is_solver = is_solver_cta && threadIdx.x == 0;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4 * my_row = A.rows + (size_t)(blockIdx.x == 0 ? 0 : blockIdx.x - 1) * kRowChunks;
    int4 * s_corr = s_dyn;
    float * s_rows = reinterpret_cast<float *>(s_dyn);

    // epochs, tracked identically by every thread of the grid
    unsigned rel = A.epoch_base; // parameter publications
    unsigned arr = A.epoch_base; // row rounds

    Solver S;
    if(is_solver)
    {
#pragma unroll
        for(int i = 0; i < 16; i++) S.resultRt[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
        for(int i = 0; i < 9; i++) S.Rcurr[i] = A.Rprev[i];
#pragma unroll
        for(int i = 0; i < 3; i++) S.tcurr[i] = A.tprev[i];
        S.last_icp_error = A.prev_icp_error; S.last_icp_count = A.prev_icp_count;
        S.last_so3_error = A.prev_so3_error; S.last_so3_count = A.prev_so3_count;
        S.last_rgb_error = A.prev_rgb_error; S.last_rgb_count = A.prev_rgb_count;
        S.so3_iterations = 0;
        S.se3_iterations[0] = S.se3_iterations[1] = S.se3_iterations[2] = 0;
    }

    int dbg_it = 0;
    auto stamp = [&](int k) {
        if(A.dbg && threadIdx.x == 0 && dbg_it < kMaxIters) A.dbg[((size_t)blockIdx.x * kMaxIters + dbg_it) * kDbgStamps + k] = clock64();
    };

    // ============================================================================================
    // SO(3) pre-alignment: RGBDOdometry.cpp:294-382 (level 2, at most 10 so3Step evaluations)
    // ============================================================================================
    if(A.so3)
    {
        const LevelArgs & L = A.lvl[2];
        So3Params P;
        P.rows = L.rows;
        P.cols = L.cols;
        P.kinv = mat_from(A.so3_kinv);
        So3State Z; // solver-private loop state
        if(is_solver)
        {
#pragma unroll
            for(int i = 0; i < 9; i++) { Z.resultR[i] = Z.lastResultR[i] = (i % 4 == 0) ? 1.0 : 0.0; Z.R_lr[i] = (i % 4 == 0) ? 1.f : 0.f; }
            Z.lastError = FLT_MAX / 2;
            Z.lastCount = FLT_MAX / 2;
        }
        UnitIter U{L.rows * L.cols, W, (int)blockIdx.x - 1, (int)lane, (int)warp};
        const int passes = U.passes();

        for(int it = 0; it <= 10; it++)
        {
            // ---- CTA 0: digest the previous evaluation (:348-380), publish the next homography ----
            if(is_solver_cta)
            {
                if(it > 0) gather_rows(A.rows, W, kSo3Chunks, arr, s_rows, s_red, s_final);
                if(is_solver)
                {
                    int done = 0;
                    if(it > 0) done = solve_so3(S, Z, s_final, it);
                    float payload[kPayload];
                    make_so3_payload(Z, payload, done, L.fx, L.fy, L.cx, L.cy, L.K_inv);
                    publish_line(ctl->line, payload, rel + 1);
                }
            }
            ++rel;
            wait_line(ctl->line, rel, s_par);
            const bool done = s_par[18] != 0.f;
            P.image_basis = mat_from(s_par);
            P.krlr = mat_from(s_par + 9);
            __syncthreads(); // s_par is rewritten by the next wait_line
            if(done) break;
            ++arr;
            if(is_solver_cta) continue;

            // ---- workers: so3Step over this CTA's pixels ----
            float acc[16];
#pragma unroll
            for(int i = 0; i < 16; i++) acc[i] = 0.f;
            for(int p = 0; p < passes; p++)
            {
                const int u = U.unit(p);
                if(u >= 0)
                {
                    const int y = u / L.cols, x = u - y * L.cols;
                    float row[4];
                    if(so3_row(P, x, y, A.so3_last, A.so3_next, L.cols, row)) accumulate_so3(acc, row);
                }
            }
            const float lane_value = warp_transpose_reduce16(acc);
            if(lane < 16) s_red[warp * 16 + lane] = lane_value;
            __syncthreads();
            if(threadIdx.x < 16)
            {
                float s = 0.f;
#pragma unroll
                for(int w = 0; w < kWarps; w++) s += s_red[w * 16 + threadIdx.x];
                s_final[threadIdx.x] = (threadIdx.x < 11) ? s : 0.f;
            }
            __syncthreads();
            publish_row(my_row, s_final, kSo3Chunks, arr);
        }
        if(is_solver)
        {
#pragma unroll
            for(int x = 0; x < 3; x++)
            {
#pragma unroll
                for(int y = 0; y < 3; y++) S.resultRt[x * 4 + y] = Z.resultR[x * 3 + y]; // :394-403
            }
        }
    }

    // ============================================================================================
    // coarse-to-fine Gauss-Newton: RGBDOdometry.cpp:405-585
    // ============================================================================================
    IcpParams IP;
    IP.Rprev_inv = mat_from(A.Rprev_inv);
    IP.tprev = make_float3(A.tprev[0], A.tprev[1], A.tprev[2]);
    IP.dist_thresh = A.dist_thresh;
    IP.angle_thresh = A.angle_thresh;

    int it_global = 0;
    bool pending = false; // a row round CTA 0 has not digested yet
    int pending_level = 0;

    // CTA 0: collect the outstanding row round, add the rows in worker order, run the reference's host step
    // (:541-583) in the solver thread
    auto solve_pending = [&]() {
        gather_rows(A.rows, W, kRowChunks, arr, s_rows, s_red, s_final);
        stamp(2);
        if(is_solver) solve_se3(S, s_final, ctl->last_S, A.icp, A.rgb, A.icp_weight, A.Rprev, A.tprev, pending_level);
        stamp(3);
    };

    for(int lv = kNumPyrs - 1; lv >= 0; lv--)
    {
        const LevelArgs & L = A.lvl[lv];
        if(L.iterations <= 0) continue;

        IP.intr = Intr{L.fx, L.fy, L.cx, L.cy};
        IP.rows = L.rows;
        IP.cols = L.cols;
        RgbResParams RP;
        RP.min_scale = L.min_scale;
        RP.max_depth_delta = A.max_depth_delta;
        RP.rows = L.rows;
        RP.cols = L.cols;
        RgbStepParams SP;
        SP.fx = L.fx; SP.fy = L.fy; SP.inv_fx = L.inv_fx; SP.inv_fy = L.inv_fy; SP.cx = L.cx; SP.cy = L.cy;
        SP.sobel_scale = A.sobel_scale;
        SP.sigma = 0.f;
        const int cols = L.cols;
        const int px = L.px;
        const int upr = cols / px; // units per row
        UnitIter U{upr * L.rows, W, (int)blockIdx.x - 1, (int)lane, (int)warp};
        const int passes = U.passes();

        float lastRGBError = FLT_MAX;      // every thread tracks it for the uniform rgb_only break (:464)
        bool first_of_level = true;

        for(int j = 0; j < L.iterations; j++)
        {
            const int cnt_slot = it_global++; // one barrier-B word per started iteration (also when it breaks)
            dbg_it = cnt_slot;
            stamp(0);
            // ---- CTA 0: finish the previous iteration, publish this one's parameters (:424-434, :480-481) ----
            if(is_solver_cta)
            {
                if(pending) solve_pending();
                if(is_solver)
                {
                    if(first_of_level) S.last_rgb_error = FLT_MAX; // :420
                    float payload[kPayload];
                    make_se3_payload(S, payload, A.rgb, L.fx, L.fy, L.cx, L.cy, L.K_inv);
                    stamp(1);
                    publish_line(ctl->line, payload, rel + 1);
                }
            }
            pending = false;
            first_of_level = false;
            ++rel;
            wait_line(ctl->line, rel, s_par);
            IP.Rcurr = mat_from(s_par);
            IP.tcurr = make_float3(s_par[9], s_par[10], s_par[11]);
            RP.krkinv = mat_from(s_par + 12);
            RP.kt = make_float3(s_par[21], s_par[22], s_par[23]);
            __syncthreads(); // s_par is rewritten by the next wait_line
            stamp(4);

            int rgbSize = 0, sigma = 0;
            unsigned long long * word = &ctl->bar_b[cnt_slot];

            // ---- workers, phase A1: photometric association -> records in shared memory, then the barrier-B arrival
            //      (arrivals | count << 8 | sigma << 32 in one 64-bit word; integer adds are exact in any order and
            //      sigma wraps mod 2^32 in the top bits exactly like the reference's int sum) ----
            if(!is_solver_cta && A.rgb)
            {
                int cnt = 0, sig = 0;
                for(int p = 0; p < passes; p++)
                {
                    const int u = U.unit(p);
                    if(u >= 0)
                    {
                        const int y = u / upr, x0 = (u - y * upr) * px;
                        if(px == 4) rgb_assoc_unit<4>(L, RP, y, x0, s_corr, p * 4, cnt, sig);
                        else rgb_assoc_unit<1>(L, RP, y, x0, s_corr, p, cnt, sig);
                    }
                    else
                    {
                        for(int k = 0; k < px; k++) s_corr[(p * px + k) * kThreads + threadIdx.x] = make_int4(-1, 0, 0, 0);
                    }
                }
                stamp(7);
                cnt = __reduce_add_sync(kFullMask, cnt);
                sig = __reduce_add_sync(kFullMask, sig);
                if(lane == 0) { s_wcnt[warp] = cnt; s_wsig[warp] = sig; }
                __syncthreads();
                if(threadIdx.x == 0)
                {
                    unsigned c = 0, s = 0;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) { c += (unsigned)s_wcnt[w]; s += (unsigned)s_wsig[w]; }
                    atomicAdd(word, 1ull | ((unsigned long long)c << 8) | ((unsigned long long)s << 32));
                }
            }
            stamp(5);

            // ---- workers, phase A2: ICP association + 29 sums (hides the barrier-B latency and the inter-CTA skew) ----
            float accI[32];
#pragma unroll
            for(int i = 0; i < 32; i++) accI[i] = 0.f;
            if(!is_solver_cta && A.icp)
            {
                for(int p = 0; p < passes; p++)
                {
                    const int u = U.unit(p);
                    if(u >= 0)
                    {
                        const int y = u / upr, x0 = (u - y * upr) * px;
                        if(px == 4) icp_unit<4>(L, IP, y, x0, accI);
                        else icp_unit<1>(L, IP, y, x0, accI);
                    }
                }
            }
            if(!is_solver_cta)
            {
                float vi = 0.f;
                if(A.icp) vi = warp_transpose_reduce32(accI);
                if(lane < 29) s_red[warp * 64 + lane] = vi;
                stamp(8);
            }

            // ---- barrier B: everybody (CTA 0 included, it needs the count for the statistics) reads the word ----
            bool level_break = false;
            if(A.rgb)
            {
                if(threadIdx.x == 0)
                {
                    unsigned long long v;
                    do
                    {
                        v = ld_relaxed64(word);
                    } while((int)(v & 0xffull) < W);
                    s_cnt = (int)((v >> 8) & 0xffffffull);
                    s_sig = (int)(unsigned)(v >> 32);
                }
                __syncthreads();
                rgbSize = s_cnt;
                sigma = s_sig;
                stamp(6);

                // RGBDOdometry.cpp:461-475 (the precedence quirk of :461 is kept)
                float sigmaVal = (float)sqrt((double)((((float)sigma / (float)rgbSize) == 0) ? 1 : rgbSize));
                const float rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
                if(A.rgb_only && rgbError > lastRGBError) level_break = true; // :464 (uniform across the grid)
                if(!level_break)
                {
                    lastRGBError = rgbError;
                    if(is_solver)
                    {
                        S.last_rgb_error = rgbError;
                        S.last_rgb_count = (float)rgbSize;
                    }
                    if(A.rgb_only) sigmaVal = -1;
                    SP.sigma = sigmaVal;
                }
            }
            else if(is_solver)
            {
                // :461-470 run even without RGB: sigma = rgbSize = 0 -> rgbError 0, count 0
                S.last_rgb_error = 0.f;
                S.last_rgb_count = 0.f;
            }
            if(level_break) break; // no row outstanding: every CTA takes the same branch

            ++arr;
            pending = true;
            pending_level = lv;
            if(is_solver_cta) continue;

            // ---- workers, phase B: photometric rows from the records in shared memory -> 29 more sums ----
            float accR[32];
#pragma unroll
            for(int i = 0; i < 32; i++) accR[i] = 0.f;
            if(A.rgb)
            {
                for(int p = 0; p < passes; p++)
                {
                    if(px == 4) rgb_rows_unit<4>(SP, s_corr, p * 4, accR);
                    else rgb_rows_unit<1>(SP, s_corr, p, accR);
                }
            }
            {
                float vr = 0.f;
                if(A.rgb) vr = warp_transpose_reduce32(accR);
                if(lane < 29) s_red[warp * 64 + 29 + lane] = vr;
                __syncthreads();
                if(threadIdx.x < kRowFloats)
                {
                    float sum = 0.f;
                    if(threadIdx.x < 58)
                    {
#pragma unroll
                        for(int w = 0; w < kWarps; w++) sum += s_red[w * 64 + threadIdx.x];
                    }
                    s_final[threadIdx.x] = sum;
                }
                __syncthreads();
                publish_row(my_row, s_final, kRowChunks, arr);
                stamp(9);
            }
        }
    }

    // ============================================================================================
    // epilogue: last solve, jump rejection (:587-591), outputs, leave the barrier-B words clean
    // ============================================================================================
    if(!pending)
    {
        // a row round without payload: tells CTA 0 that every worker is past its last barrier-B wait
        ++arr;
        if(!is_solver_cta)
        {
            __syncthreads();
            publish_row(my_row, s_final, kSo3Chunks, arr);
        }
    }
    if(is_solver_cta)
    {
        if(pending) solve_pending();
        else gather_rows(A.rows, W, kSo3Chunks, arr, s_rows, s_red, s_final);
    }
    if(is_solver)
    {
        if(A.rgb)
        {
            const float d[3] = {S.tcurr[0] - A.tprev[0], S.tcurr[1] - A.tprev[1], S.tcurr[2] - A.tprev[2]};
            if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
            {
#pragma unroll
                for(int i = 0; i < 9; i++) S.Rcurr[i] = A.Rprev[i];
#pragma unroll
                for(int i = 0; i < 3; i++) S.tcurr[i] = A.tprev[i];
            }
        }
        TrackOutput * out = A.out;
#pragma unroll
        for(int i = 0; i < 3; i++) out->trans[i] = S.tcurr[i];
#pragma unroll
        for(int i = 0; i < 9; i++) out->rot[i] = S.Rcurr[i];
        out->st.last_icp_error = S.last_icp_error; out->st.last_icp_count = S.last_icp_count;
        out->st.last_rgb_error = S.last_rgb_error; out->st.last_rgb_count = S.last_rgb_count;
        out->st.last_so3_error = S.last_so3_error; out->st.last_so3_count = S.last_so3_count;
        out->st.so3_iterations = S.so3_iterations;
#pragma unroll
        for(int i = 0; i < 3; i++) out->st.se3_iterations[i] = S.se3_iterations[i];
        if(S.se3_iterations[0] + S.se3_iterations[1] + S.se3_iterations[2] > 0)
        {
            // lastA / lastb of the last solve (reduce.cu:475-486 unpack order)
#pragma unroll
            for(int i = 0; i < 6; i++)
            {
#pragma unroll
                for(int j = i; j < 7; j++)
                {
                    const double v = ctl->last_S[hm::acc_index(i, j)];
                    if(j == 6) out->st.last_b[i] = v;
                    else out->st.last_A[j * 6 + i] = out->st.last_A[i * 6 + j] = v;
                }
            }
            out->status = 1;
        }
        else
            out->status = 2; // no solve ran: lastA / lastb keep their previous values (host side)
        __threadfence_system();
        // every worker has published its last row and touches the control block no more
        for(int i = 0; i < kMaxIters; i++) ctl->bar_b[i] = 0ull;
    }
}

struct DeviceTrack
{
    long long * dbg;      // device, max_grid * kMaxIters * kDbgStamps stamps (EF_TRACK_TIMING=1)
    double dbg_acc[kMaxIters][kDbgStamps];
    double wrk_mean[kMaxIters][5], wrk_max[kMaxIters][5]; // worker phase durations, mean / max over the worker CTAs
    long long * dbg_host;
    int dbg_grid;
    long long dbg_n;
    TrackCtl * ctl;
    uint4 * rows;
    TrackOutput * out; // pinned
    int grid;
    int px[kNumPyrs];
    size_t smem_bytes;
    unsigned launch_seq;
};

} // namespace

// (re)derive the launch geometry for a grid of `grid` CTAs: pixels per thread unit and passes per level ->
// shared-memory records per thread.  A handle normally owns every SM; EF_OPT_GRID_CTAS lets several handles
// share the GPU (e.g. two sequences tracked concurrently on 74 SMs each).
int device_track_configure(ef_tracker * t, int grid)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    const int max_grid = t->num_sms < 255 ? t->num_sms : 255; // arrivals live in 8 bits of the barrier-B word
    if(grid <= 0 || grid > max_grid) grid = max_grid;
    if(grid < 2) grid = 2;
    d->grid = grid;
    const int W = d->grid - 1;
    int max_records = 1;
    for(int i = 0; i < kNumPyrs; i++)
    {
        const int npix = t->dims[i].rows * t->dims[i].cols;
        const int px = ((npix + W - 1) / W > kThreads && (t->dims[i].cols % 4) == 0) ? 4 : 1;
        d->px[i] = px;
        const int chunks = (npix / px + 31) / 32;
        const int per_worker = (chunks + W - 1) / W;
        const int passes = (per_worker + kWarps - 1) / kWarps;
        if(passes * px > max_records) max_records = passes * px;
    }
    d->smem_bytes = (size_t)max_records * kThreads * sizeof(int4);
    const size_t rows_smem = (size_t)W * kRowFloats * sizeof(float);
    if(rows_smem > d->smem_bytes) d->smem_bytes = rows_smem;
    return EF_OK;
}

int device_track_init(ef_tracker * t)
{
    DeviceTrack * d = new DeviceTrack();
    memset(d, 0, sizeof(*d));
    t->track_state = d;
    device_track_configure(t, 0);
    const int max_grid = t->num_sms < 255 ? t->num_sms : 255;
    d->launch_seq = 0;
    cudaError_t e = cudaMalloc((void **)&d->ctl, sizeof(TrackCtl));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->ctl, 0, sizeof(TrackCtl), t->stream);
    if(e == cudaSuccess) e = cudaMalloc((void **)&d->rows, (size_t)max_grid * kRowChunks * sizeof(uint4));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->rows, 0, (size_t)max_grid * kRowChunks * sizeof(uint4), t->stream);
    if(e == cudaSuccess) e = cudaHostAlloc((void **)&d->out, sizeof(TrackOutput), cudaHostAllocMapped);
    const char * env = getenv("EF_TRACK_TIMING");
    if(e == cudaSuccess && env && env[0] == '1')
    {
        d->dbg_grid = max_grid;
        d->dbg_host = (long long *)malloc(sizeof(long long) * max_grid * kMaxIters * kDbgStamps);
        e = cudaMalloc((void **)&d->dbg, sizeof(long long) * max_grid * kMaxIters * kDbgStamps);
        if(e == cudaSuccess) e = cudaMemsetAsync(d->dbg, 0, sizeof(long long) * max_grid * kMaxIters * kDbgStamps, t->stream);
    }
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_track, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if(e != cudaSuccess)
    {
        device_track_destroy(t);
        return (int)e;
    }
    t->h_track_out = d->out;
    return EF_OK;
}

void device_track_destroy(ef_tracker * t)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return;
    if(d->dbg)
    {
        if(d->dbg_n > 0)
        {
            // stamps (cycles of CTA 0's SM): 0 loop top | 1 arrivals seen | 2 final reduce | 3 solve | 4 params published
            // + loaded | 5 phase A | 6 barrier B | 7 phase B | 8 reduce + arrive
            fprintf(stderr, "[ef_track timing] avg cycles of CTA 0 per SE3 iteration over %lld calls\n", d->dbg_n);
            for(int it = 0; it < kMaxIters; it++)
            {
                const double * a = d->dbg_acc[it];
                if(a[4] == 0) continue;
                const double n = (double)d->dbg_n;
                const bool first = a[2] == 0;
                fprintf(stderr, "  it %2d: gather %7.0f solve %7.0f payload %7.0f publish+detect %7.0f | to-barB-seen %7.0f | total %8.0f\n", it,
                        first ? 0.0 : (a[2] - a[0]) / n, first ? 0.0 : (a[3] - a[2]) / n, (a[1] - (first ? a[0] : a[3])) / n, (a[4] - a[1]) / n,
                        (a[6] - a[4]) / n, (a[6] - a[0]) / n);
            }
            // worker CTAs: params wait (0->4) | photometric association + CTA sync (4->5) | ICP + warp reduce (5->8) |
            // barrier-B wait (8->6) | photometric rows + CTA reduce + publish (6->9)
            fprintf(stderr, "[ef_track timing] worker CTAs, cycles mean/max over workers\n");
            for(int it = 0; it < kMaxIters; it++)
            {
                if(d->dbg_acc[it][4] == 0) continue;
                const double n = (double)d->dbg_n;
                const double * m = d->wrk_mean[it];
                const double * x = d->wrk_max[it];
                fprintf(stderr, "  it %2d: wait-params %6.0f/%6.0f  rgb-assoc %6.0f/%6.0f  icp %6.0f/%6.0f  wait-barB %6.0f/%6.0f  rgb-rows+publish %6.0f/%6.0f\n", it,
                        m[0] / n, x[0] / n, m[1] / n, x[1] / n, m[2] / n, x[2] / n, m[3] / n, x[3] / n, m[4] / n, x[4] / n);
            }
        }
        cudaFree(d->dbg);
        free(d->dbg_host);
    }
    if(d->ctl) cudaFree(d->ctl);
    if(d->rows) cudaFree(d->rows);
    if(d->out) cudaFreeHost(d->out);
    delete d;
    t->track_state = nullptr;
}

int device_track_launch(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    if(d->smem_bytes > 200 * 1024 || (size_t)t->width * t->height >= (1u << 24) || t->width >= 32768 || t->height >= 32768)
    {
        t->err = "image too large for the shared-memory correspondence store of EF_SOLVE_DEVICE";
        return EF_ERR_UNSUPPORTED;
    }
    TrackArgs A;
    memset(&A, 0, sizeof(A));
    const int iterations[kNumPyrs] = {fast_odom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0}; // :384-386
    for(int i = 0; i < kNumPyrs; i++)
    {
        LevelArgs & L = A.lvl[i];
        L.vc = t->vmap_curr[i]; L.nc = t->nmap_curr[i]; L.vp = t->vmap_g_prev[i]; L.np = t->nmap_g_prev[i];
        L.last_depth = t->last_depth[i]; L.next_depth = t->next_depth[i];
        L.last_image = t->last_image[i]; L.next_image = t->next_image[i];
        L.dIdx = t->dIdx[i]; L.dIdy = t->dIdy[i];
        L.rows = t->dims[i].rows; L.cols = t->dims[i].cols;
        const int div = 1 << i;
        L.fx = t->fx / div; L.fy = t->fy / div; L.cx = t->cx / div; L.cy = t->cy / div;
        L.inv_fx = 1.0f / L.fx; L.inv_fy = 1.0f / L.fy;
        L.min_scale = (float)(pow(t->min_grad[i], 2.0) / pow(t->sobel_scale, 2.0)); // :442
        L.iterations = iterations[i];
        L.px = d->px[i];
        const double K[9] = {L.fx, 0, L.cx, 0, L.fy, L.cy, 0, 0, 1};
        hm::inverse33(K, L.K_inv);
    }
    A.so3_last = t->last_next_image[2];
    A.so3_next = t->next_image[2];
    for(int i = 0; i < 9; i++) A.so3_kinv[i] = (float)A.lvl[2].K_inv[i];
    memcpy(A.Rprev, rot, sizeof(A.Rprev));
    memcpy(A.tprev, trans, sizeof(A.tprev));
    hm::inverse33(A.Rprev, A.Rprev_inv); // :388
    A.dist_thresh = t->dist_thresh; A.angle_thresh = t->angle_thresh;
    A.max_depth_delta = t->max_depth_delta_rgb; A.sobel_scale = t->sobel_scale; A.icp_weight = icp_weight;
    A.icp = (!rgb_only && icp_weight > 0) ? 1 : 0;
    A.rgb = (rgb_only || icp_weight < 100) ? 1 : 0;
    A.rgb_only = rgb_only ? 1 : 0;
    A.so3 = so3 ? 1 : 0;
    A.prev_icp_error = t->st.last_icp_error; A.prev_icp_count = t->st.last_icp_count;
    A.prev_so3_error = t->st.last_so3_error; A.prev_so3_count = t->st.last_so3_count;
    A.prev_rgb_error = t->st.last_rgb_error; A.prev_rgb_count = t->st.last_rgb_count;
    // launch-unique flag epochs: nothing in the control block or the rows needs resetting between launches.  After
    // 2^24 launches the sequence wraps; stale flags are wiped then.
    d->launch_seq++;
    if((d->launch_seq & 0xffffffu) == 0)
    {
        d->launch_seq = 1;
        cudaMemsetAsync(d->rows, 0, (size_t)(t->num_sms < 255 ? t->num_sms : 255) * kRowChunks * sizeof(uint4), t->stream);
        cudaMemsetAsync(d->ctl, 0, sizeof(TrackCtl), t->stream);
    }
    A.epoch_base = d->launch_seq << 8;
    A.ctl = d->ctl;
    A.rows = d->rows;
    A.out = d->out; // UVA: pinned + mapped host memory is addressable from the device
    A.dbg = d->dbg;

    d->out->status = 0;
    void * args[] = {&A};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_track, dim3(d->grid), dim3(kThreads), args, d->smem_bytes, t->stream);
    t->launches++;
    if(e != cudaSuccess)
    {
        t->err = std::string("cudaLaunchCooperativeKernel: ") + cudaGetErrorString(e);
        return (int)e;
    }
    return EF_OK;
}

int device_track_finish(ef_tracker * t, float * trans, float * rot)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    cudaError_t e = cudaStreamSynchronize(t->stream);
    if(e != cudaSuccess)
    {
        t->err = std::string("track kernel: ") + cudaGetErrorString(e);
        return (int)e;
    }
    if(d->out->status != 1 && d->out->status != 2)
    {
        t->err = "track kernel produced no result";
        return EF_ERR_BAD_STATE;
    }
    memcpy(trans, d->out->trans, sizeof(d->out->trans));
    memcpy(rot, d->out->rot, sizeof(d->out->rot));
    const ef_track_stats & o = d->out->st;
    t->st.last_icp_error = o.last_icp_error; t->st.last_icp_count = o.last_icp_count;
    t->st.last_rgb_error = o.last_rgb_error; t->st.last_rgb_count = o.last_rgb_count;
    t->st.last_so3_error = o.last_so3_error; t->st.last_so3_count = o.last_so3_count;
    t->st.so3_iterations = o.so3_iterations;
    for(int i = 0; i < 3; i++) t->st.se3_iterations[i] = o.se3_iterations[i];
    if(d->out->status == 1)
    {
        memcpy(t->st.last_A, o.last_A, sizeof(o.last_A));
        memcpy(t->st.last_b, o.last_b, sizeof(o.last_b));
    }
    if(d->dbg)
    {
        long long * h = d->dbg_host;
        const size_t bytes = sizeof(long long) * d->dbg_grid * kMaxIters * kDbgStamps;
        if(cudaMemcpy(h, d->dbg, bytes, cudaMemcpyDeviceToHost) == cudaSuccess)
        {
            for(int it = 0; it < kMaxIters; it++)
            {
                for(int k = 0; k < kDbgStamps; k++) d->dbg_acc[it][k] += (double)h[it * kDbgStamps + k];
                static const int from[5] = {0, 4, 5, 8, 6}, to[5] = {4, 5, 8, 6, 9};
                for(int ph = 0; ph < 5; ph++)
                {
                    double sum = 0, mx = 0;
                    int cnt = 0;
                    for(int c = 1; c < d->grid; c++)
                    {
                        const long long * w = h + ((size_t)c * kMaxIters + it) * kDbgStamps;
                        if(w[from[ph]] == 0 || w[to[ph]] == 0) continue;
                        const double v = (double)(w[to[ph]] - w[from[ph]]);
                        sum += v;
                        if(v > mx) mx = v;
                        cnt++;
                    }
                    if(cnt)
                    {
                        d->wrk_mean[it][ph] += sum / cnt;
                        d->wrk_max[it][ph] += mx;
                    }
                }
            }
            d->dbg_n++;
            cudaMemset(d->dbg, 0, bytes);
        }
    }
    return EF_OK;
}

} // namespace ef
