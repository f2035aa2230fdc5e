// ef_track_kernel.cu -- EF_SOLVE_DEVICE: the whole of RGBDOdometry::getIncrementalTransformation
// (elasticfusionpublic/Core/src/Utils/RGBDOdometry.cpp:292-591) as ONE persistent cooperative kernel.
//
// Why: at 640x480 a Gauss-Newton iteration touches ~19 MB that already sits in the 126 MB L2, i.e. a
// couple of microseconds of memory time, while the reference pays 3-4 launches, 3-4 device syncs and
// blocking D2H copies per iteration (57+ host round trips per frame).  Here one CTA per SM stays
// resident for the whole solve:
//
//   per iteration   phase A  every thread: ICP association + 29 fp32 sums for its pixels (registers),
//                            photometric association -> correspondence records in SHARED memory,
//                            {count, sum diff^2} by integer atomics
//                   barrier  (only when RGB is on: sigma depends on the global count)
//                   phase B  photometric rows from the records in shared memory -> 29 more sums
//                   reduce   transpose-reduce butterfly -> per-CTA 64-float partial row in global memory
//                   solve    CTA 0 waits for all arrivals, adds the rows in CTA order (deterministic),
//                            one thread runs the reference's host step in double (6x6 LDLT, exp map,
//                            pose composition: ef_hostmath.h, the same code the host path runs) and
//                            publishes the next iteration's parameters; everyone else spins on a flag.
//
// No DataTerm image, no point cloud, no reduceSum launch, no host involvement until the final pose is
// stored straight into pinned host memory.  Grid barriers are hand-rolled (red.release / ld.acquire on
// monotonically increasing counters); co-residency is guaranteed by a cooperative launch with
// gridDim = number of SMs.
#include <cooperative_groups.h>
#include <float.h>
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

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxIters = 32; // SE3 iterations per call (19 in the reference schedule)

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
};

// parameters the solver publishes for the next phase (read by every CTA after the barrier)
struct TrackParams
{
    float Rcurr[9], tcurr[3];   // ICP
    float krkinv[9], kt[3];     // RGB warp
    float H[9], krlr[9];        // SO3 homography and K*R
    int so3_done;
    int pad[3];
};

struct TrackCtl
{
    unsigned arrive;  unsigned pad0[31];          // solve barrier: arrivals (monotonic within a launch)
    unsigned release; unsigned pad1[31];          // solve barrier: epochs released by CTA 0
    unsigned arrive_b; unsigned pad2[31];         // phase A -> B barrier arrivals
    int rgb_cnt[kMaxIters][2];                    // per SE3 iteration {count, sum (int)(diff^2)}
    TrackParams params;
};

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
    TrackCtl * ctl;
    float * partials;                             // gridDim.x * 64 floats
    TrackOutput * out;
    int corr_slots;                               // int4 records per thread in dynamic shared memory
};

// ---- memory-ordering primitives ----
__device__ __forceinline__ unsigned ld_acquire(const unsigned * p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned * p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all threads of the CTA call; thread 0 signals arrival after the CTA's writes are visible
__device__ __forceinline__ void cta_arrive(unsigned * counter)
{
    __syncthreads();
    if(threadIdx.x == 0)
    {
        __threadfence();
        atomicAdd(counter, 1u);
    }
}

// all threads call; returns when *counter >= target (acquire)
__device__ __forceinline__ void cta_wait_ge(const unsigned * counter, unsigned target)
{
    if(threadIdx.x == 0)
    {
        while(ld_acquire(counter) < target) { }
        __threadfence();
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
    ef_track_stats st;
};

// CTA 0 only, after all arrivals: add the partial rows in CTA-index order -> s_final[64]
__device__ __forceinline__ void cta0_final_reduce(const float * __restrict__ partials, float * s_red, float * s_final)
{
    const int slot = threadIdx.x & 63, part = threadIdx.x >> 6; // 8 parts of 64 lanes
    float s = 0.f;
    for(unsigned b = part; b < gridDim.x; b += kThreads / 64) s += __ldcg(partials + b * 64 + slot);
    __syncthreads();
    s_red[part * 64 + slot] = s;
    __syncthreads();
    if(threadIdx.x < 64)
    {
        float tot = 0.f;
#pragma unroll
        for(int p = 0; p < kThreads / 64; p++) tot += s_red[p * 64 + threadIdx.x];
        s_final[threadIdx.x] = tot;
    }
    __syncthreads();
}


// RGBDOdometry.cpp:515-516, :541-583 -- runs in ONE thread (thread 0 of CTA 0).  Kept out of line so
// its double-precision register needs do not inflate the per-pixel phases of the kernel.
__device__ __noinline__ void solve_se3(Solver & S, const float * s_final, int icp, int rgb, float icp_weight, const float * Rprev,
                                       const float * tprev, int level)
{
    double A_icp[36], b_icp[6], A_rgb[36], b_rgb[6];
    if(icp)
    {
        float residual[2];
        hm::unpack_se3(s_final, A_icp, b_icp, residual);
        S.st.last_icp_error = sqrtf(residual[0]) / residual[1]; // :515-516
        S.st.last_icp_count = residual[1];
    }
    if(rgb) hm::unpack_se3(s_final + 32, A_rgb, b_rgb, (float *)nullptr);
    double * lastA = S.st.last_A, * lastb = S.st.last_b, result[6];
    if(icp && rgb) // :547-553
    {
        const double w = icp_weight;
        for(int k = 0; k < 36; k++) lastA[k] = A_rgb[k] + w * w * A_icp[k];
        for(int k = 0; k < 6; k++) lastb[k] = b_rgb[k] + w * b_icp[k];
    }
    else if(icp)
    {
        for(int k = 0; k < 36; k++) lastA[k] = A_icp[k];
        for(int k = 0; k < 6; k++) lastb[k] = b_icp[k];
    }
    else
    {
        for(int k = 0; k < 36; k++) lastA[k] = A_rgb[k];
        for(int k = 0; k < 6; k++) lastb[k] = b_rgb[k];
    }
    hm::ldlt_solve<double, 6>(lastA, lastb, result);
    S.st.se3_iterations[level]++;
    hm::update_se3(S.resultRt, result);                           // :573
    hm::compose_pose(S.resultRt, Rprev, tprev, S.Rcurr, S.tcurr); // :575-583
}

// :424-434 -- publish the parameters of the next SE3 iteration (solver thread)
__device__ __noinline__ void publish_se3(const Solver & S, TrackParams * p, int rgb, float fx, float fy, float cx, float cy)
{
    for(int i = 0; i < 9; i++) p->Rcurr[i] = S.Rcurr[i];
    for(int i = 0; i < 3; i++) p->tcurr[i] = S.tcurr[i];
    if(rgb)
    {
        const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
        double K_inv[9];
        hm::inverse33(K, K_inv);
        hm::rgb_warp_params(S.resultRt, K, K_inv, p->krkinv, p->kt);
    }
}

// :348-380 -- digest one so3Step evaluation; returns done
struct So3State
{
    double resultR[9], lastResultR[9];
    float R_lr[9], lastError, lastCount;
};

__device__ __noinline__ int solve_so3(Solver & S, So3State & Z, const float * s_final, int it)
{
    int done = 0;
    float jtj[9], jtr[3], residual[2];
    hm::unpack_so3(s_final, jtj, jtr, residual);
    S.st.so3_iterations++;
    S.st.last_so3_error = sqrtf(residual[0]) / residual[1];                           // :348
    S.st.last_so3_count = residual[1];
    if(S.st.last_so3_error < Z.lastError && Z.lastCount == S.st.last_so3_count) done = 1; // :352
    else if(S.st.last_so3_error > Z.lastError + 0.001)                                 // :356
    {
        S.st.last_so3_error = Z.lastError;
        S.st.last_so3_count = Z.lastCount;
        for(int i = 0; i < 9; i++) Z.resultR[i] = Z.lastResultR[i];
        done = 1;
    }
    if(!done)
    {
        Z.lastError = S.st.last_so3_error;
        Z.lastCount = S.st.last_so3_count;
        for(int i = 0; i < 9; i++) Z.lastResultR[i] = Z.resultR[i];
        float delta[3];
        hm::ldlt_solve<float, 3>(jtj, jtr, delta);                                     // :368
        const double dd[3] = {delta[0], delta[1], delta[2]};
        double rotUpdate[9];
        hm::rodrigues(dd, rotUpdate);
        float ru[9];
        for(int i = 0; i < 9; i++) ru[i] = (float)rotUpdate[i];
        hm::mul33(ru, Z.R_lr, Z.R_lr);                                                  // :372
        for(int i = 0; i < 9; i++) Z.resultR[i] = Z.R_lr[i];
        if(it == 10) done = 1; // ten evaluations made
    }
    return done;
}

// :318-329 -- homography K R K^-1 and K R for the next so3Step
__device__ __noinline__ void publish_so3(const So3State & Z, TrackParams * p, int done, float fx, float fy, float cx, float cy)
{
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    double K_inv[9], KR[9], H[9];
    hm::inverse33(K, K_inv);
    hm::mul33(K, Z.resultR, KR);
    hm::mul33(KR, K_inv, H);
    for(int i = 0; i < 9; i++) { p->H[i] = (float)H[i]; p->krlr[i] = (float)KR[i]; }
    p->so3_done = done;
}

__global__ void __launch_bounds__(kThreads, 1) k_track(const TrackArgs A)
{
    extern __shared__ int4 s_corr[];            // [corr_slots][kThreads]
    __shared__ float s_red[kWarps * 64];
    __shared__ float s_final[64];
    __shared__ TrackParams s_par;
    __shared__ int s_cnt[kWarps], s_sig[kWarps];

    TrackCtl * ctl = A.ctl;
    const unsigned grid = gridDim.x;
    const bool is_solver_cta = (blockIdx.x == 0);
    const bool is_solver = is_solver_cta && threadIdx.x == 0;
    float * my_row = A.partials + (size_t)blockIdx.x * 64;

    // barrier bookkeeping, tracked identically by every thread of the grid
    unsigned rel = 0;   // parameter publications (release epochs)
    unsigned arr = 0;   // completed arrival rounds on ctl->arrive
    unsigned arr_b = 0; // completed arrival rounds on ctl->arrive_b

    Solver S;
    if(is_solver)
    {
        for(int i = 0; i < 16; i++) S.resultRt[i] = (i % 5 == 0) ? 1.0 : 0.0;
        for(int i = 0; i < 9; i++) S.Rcurr[i] = A.Rprev[i];
        for(int i = 0; i < 3; i++) S.tcurr[i] = A.tprev[i];
        memset(&S.st, 0, sizeof(S.st));
        S.st.last_icp_error = A.prev_icp_error; S.st.last_icp_count = A.prev_icp_count;
        S.st.last_so3_error = A.prev_so3_error; S.st.last_so3_count = A.prev_so3_count;
        S.st.last_rgb_error = A.prev_rgb_error; S.st.last_rgb_count = A.prev_rgb_count;
    }

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
            for(int i = 0; i < 9; i++) { Z.resultR[i] = Z.lastResultR[i] = (i % 4 == 0) ? 1.0 : 0.0; Z.R_lr[i] = (i % 4 == 0) ? 1.f : 0.f; }
            Z.lastError = FLT_MAX / 2;
            Z.lastCount = FLT_MAX / 2;
        }

        const int npix = L.rows * L.cols;
        const int per_cta = (npix + grid - 1) / grid;
        const int p_begin = min(npix, (int)blockIdx.x * per_cta), p_end = min(npix, p_begin + per_cta);

        for(int it = 0; it <= 10; it++)
        {
            // ---- CTA 0: digest the previous evaluation (:348-380), publish the next homography ----
            if(is_solver_cta)
            {
                int done = 0;
                if(it > 0)
                {
                    cta_wait_ge(&ctl->arrive, arr * grid);
                    const int slot = threadIdx.x & 15, part = threadIdx.x >> 4; // 32 parts
                    float s = 0.f;
                    for(unsigned b = part; b < grid; b += kThreads / 16) s += __ldcg(A.partials + b * 64 + slot);
                    s_red[part * 16 + slot] = s;
                    __syncthreads();
                    if(threadIdx.x < 16)
                    {
                        float tot = 0.f;
                        for(int p = 0; p < kThreads / 16; p++) tot += s_red[p * 16 + threadIdx.x];
                        s_final[threadIdx.x] = tot;
                    }
                    __syncthreads();
                }
                if(is_solver)
                {
                    if(it > 0) done = solve_so3(S, Z, s_final, it);
                    publish_so3(Z, &ctl->params, done, L.fx, L.fy, L.cx, L.cy);
                    __threadfence();
                    st_release(&ctl->release, rel + 1);
                }
            }
            ++rel;
            cta_wait_ge(&ctl->release, rel);
            if(threadIdx.x < 18) (&s_par.H[0])[threadIdx.x] = __ldcg(&ctl->params.H[0] + threadIdx.x); // H then krlr are contiguous
            if(threadIdx.x == 32) s_par.so3_done = __ldcg(&ctl->params.so3_done);
            __syncthreads();
            const int done = s_par.so3_done;
            P.image_basis = mat_from(s_par.H);
            P.krlr = mat_from(s_par.krlr);
            __syncthreads();
            if(done) break;

            // ---- so3Step over this CTA's pixels ----
            float acc[16];
#pragma unroll
            for(int i = 0; i < 16; i++) acc[i] = 0.f;
            for(int k = p_begin + threadIdx.x; k < p_end; k += kThreads)
            {
                const int y = k / L.cols, x = k - y * L.cols;
                float row[4];
                if(so3_row(P, x, y, A.so3_last, A.so3_next, L.cols, row)) accumulate_so3(acc, row);
            }
            const float lane_value = warp_transpose_reduce16(acc);
            {
                const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
                if(lane < 16) s_red[warp * 16 + lane] = lane_value;
                __syncthreads();
                if(threadIdx.x < 16)
                {
                    float s = 0.f;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) s += s_red[w * 16 + threadIdx.x];
                    my_row[threadIdx.x] = s;
                }
            }
            cta_arrive(&ctl->arrive);
            ++arr;
        }
        if(is_solver)
            for(int x = 0; x < 3; x++)
                for(int y = 0; y < 3; y++) S.resultRt[x * 4 + y] = Z.resultR[x * 3 + y]; // :394-403
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
    bool pending = false; // an arrival round whose sums CTA 0 has not digested yet
    int pending_level = 0;

    // CTA 0: wait for the outstanding arrival round, add the partial rows in CTA order, run the
    // reference's host step (:541-583) in the solver thread
    auto solve_pending = [&]() {
        cta_wait_ge(&ctl->arrive, arr * grid);
        cta0_final_reduce(A.partials, s_red, s_final);
        if(is_solver) solve_se3(S, s_final, A.icp, A.rgb, A.icp_weight, A.Rprev, A.tprev, pending_level);
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
        const Map3 vc{L.vc, L.cols, L.rows}, nc{L.nc, L.cols, L.rows}, vp{L.vp, L.cols, L.rows}, np{L.np, L.cols, L.rows};

        float lastRGBError = FLT_MAX;      // every thread tracks it for the uniform rgb_only break (:464)
        bool first_of_level = true;

        // this CTA's contiguous run of 4-pixel groups
        const int gpr = L.cols >> 2;
        const int ngroups = gpr * L.rows;
        const int per_cta = (ngroups + grid - 1) / grid;
        const int g_begin = min(ngroups, (int)blockIdx.x * per_cta), g_end = min(ngroups, g_begin + per_cta);

        for(int j = 0; j < L.iterations; j++)
        {
            const int cnt_slot = it_global++; // one {count, sigma} slot per started iteration (also when it breaks)
            // ---- CTA 0: finish the previous iteration, publish this one's parameters (:424-434, :480-481) ----
            if(is_solver_cta)
            {
                if(pending) solve_pending();
                if(is_solver)
                {
                    if(first_of_level) S.st.last_rgb_error = FLT_MAX; // :420
                    publish_se3(S, &ctl->params, A.rgb, L.fx, L.fy, L.cx, L.cy);
                    __threadfence();
                    st_release(&ctl->release, rel + 1);
                }
            }
            pending = false;
            first_of_level = false;
            ++rel;
            cta_wait_ge(&ctl->release, rel);
            if(threadIdx.x < 24) (&s_par.Rcurr[0])[threadIdx.x] = __ldcg(&ctl->params.Rcurr[0] + threadIdx.x); // Rcurr,tcurr,krkinv,kt
            __syncthreads();
            IP.Rcurr = mat_from(s_par.Rcurr);
            IP.tcurr = make_float3(s_par.tcurr[0], s_par.tcurr[1], s_par.tcurr[2]);
            RP.krkinv = mat_from(s_par.krkinv);
            RP.kt = make_float3(s_par.kt[0], s_par.kt[1], s_par.kt[2]);

            float accI[32];
#pragma unroll
            for(int i = 0; i < 32; i++) accI[i] = 0.f;
            int cnt = 0, sig = 0;

            // ---- phase A ----
            int slot = 0;
            for(int g = g_begin + threadIdx.x; g < g_end; g += kThreads, slot++)
            {
                const int y = g / gpr;
                const int x0 = (g - y * gpr) << 2;
                const size_t o = (size_t)y * L.cols + x0;
                if(A.icp)
                {
                    const float4 a = *reinterpret_cast<const float4 *>(vc.row(0, y) + x0);
                    const float4 b = *reinterpret_cast<const float4 *>(vc.row(1, y) + x0);
                    const float4 c = *reinterpret_cast<const float4 *>(vc.row(2, y) + x0);
                    const float4 d = *reinterpret_cast<const float4 *>(nc.row(0, y) + x0);
                    const float4 e = *reinterpret_cast<const float4 *>(nc.row(1, y) + x0);
                    const float4 f = *reinterpret_cast<const float4 *>(nc.row(2, y) + x0);
                    const float vx[4] = {a.x, a.y, a.z, a.w}, vy[4] = {b.x, b.y, b.z, b.w}, vz[4] = {c.x, c.y, c.z, c.w};
                    const float nx[4] = {d.x, d.y, d.z, d.w}, ny[4] = {e.x, e.y, e.z, e.w}, nz[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                    for(int k = 0; k < 4; k++)
                    {
                        float row[7];
                        if(icp_row(IP, make_float3(vx[k], vy[k], vz[k]), make_float3(nx[k], ny[k], nz[k]), vp, np, row)) accumulate_se3(accI, row);
                    }
                }
                if(A.rgb)
                {
                    const short4 gx4 = *reinterpret_cast<const short4 *>(L.dIdx + o);
                    const short4 gy4 = *reinterpret_cast<const short4 *>(L.dIdy + o);
                    const float4 d14 = *reinterpret_cast<const float4 *>(L.next_depth + o);
                    const short gxs[4] = {gx4.x, gx4.y, gx4.z, gx4.w}, gys[4] = {gy4.x, gy4.y, gy4.z, gy4.w};
                    const float d1s[4] = {d14.x, d14.y, d14.z, d14.w};
#pragma unroll
                    for(int k = 0; k < 4; k++)
                    {
                        int u0 = 0, v0 = 0;
                        float diff = 0.f, d0 = 0.f;
                        const bool ok = rgb_residual_px(RP, x0 + k, y, gxs[k], gys[k], d1s[k], L.next_image, L.cols, L.last_image, L.last_depth,
                                                        L.cols, u0, v0, diff, d0);
                        int4 rec;
                        rec.x = ok ? ((u0 & 0xffff) | (v0 << 16)) : -1;
                        rec.y = __float_as_int(diff);
                        rec.z = __float_as_int(d0);
                        rec.w = ((int)gxs[k] & 0xffff) | ((int)gys[k] << 16);
                        s_corr[(slot * 4 + k) * kThreads + threadIdx.x] = rec;
                        if(ok)
                        {
                            cnt += 1;
                            sig += (int)(diff * diff); // reduce.cu:830
                        }
                    }
                }
            }

            // ICP sums leave the registers before the photometric phase needs its own 29
            {
                const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
                float vi = 0.f;
                if(A.icp) vi = warp_transpose_reduce32(accI);
                s_red[warp * 64 + lane] = vi;
            }
            float accR[32];
#pragma unroll
            for(int i = 0; i < 32; i++) accR[i] = 0.f;

            bool level_break = false;
            if(A.rgb)
            {
                // ---- {count, sigma}: integer adds are exact in any order ----
                cnt = __reduce_add_sync(kFullMask, cnt);
                sig = __reduce_add_sync(kFullMask, sig);
                const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
                if(lane == 0) { s_cnt[warp] = cnt; s_sig[warp] = sig; }
                __syncthreads();
                if(threadIdx.x == 0)
                {
                    int c = 0, s = 0;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) { c += s_cnt[w]; s += s_sig[w]; }
                    if(c) atomicAdd(&ctl->rgb_cnt[cnt_slot][0], c);
                    if(s) atomicAdd(&ctl->rgb_cnt[cnt_slot][1], s);
                }
                cta_arrive(&ctl->arrive_b);
                ++arr_b;
                cta_wait_ge(&ctl->arrive_b, arr_b * grid);
                if(threadIdx.x == 0)
                {
                    s_cnt[0] = __ldcg(&ctl->rgb_cnt[cnt_slot][0]);
                    s_sig[0] = __ldcg(&ctl->rgb_cnt[cnt_slot][1]);
                }
                __syncthreads();
                const int rgbSize = s_cnt[0], sigma = s_sig[0];
                __syncthreads();

                // RGBDOdometry.cpp:461-475 (the precedence quirk of :461 is kept)
                float sigmaVal = (float)sqrt((double)((((float)sigma / (float)rgbSize) == 0) ? 1 : rgbSize));
                const float rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
                if(A.rgb_only && rgbError > lastRGBError) level_break = true; // :464 (uniform across the grid)
                if(!level_break)
                {
                    lastRGBError = rgbError;
                    if(is_solver)
                    {
                        S.st.last_rgb_error = rgbError;
                        S.st.last_rgb_count = (float)rgbSize;
                    }
                    if(A.rgb_only) sigmaVal = -1;
                    SP.sigma = sigmaVal;

                    // ---- phase B: photometric rows from the records in shared memory ----
                    int sl = 0;
                    for(int g = g_begin + threadIdx.x; g < g_end; g += kThreads, sl++)
                    {
#pragma unroll
                        for(int k = 0; k < 4; k++)
                        {
                            const int4 rec = s_corr[(sl * 4 + k) * kThreads + threadIdx.x];
                            if(rec.x != -1)
                            {
                                const int u0 = rec.x & 0xffff, v0 = rec.x >> 16;
                                const float Z = __int_as_float(rec.z);
                                const float3 cp = project_point(u0, v0, Z, SP.inv_fx, SP.inv_fy, SP.cx, SP.cy);
                                float row[7];
                                rgb_row(SP, __int_as_float(rec.y), cp.x, cp.y, cp.z, (short)(rec.w & 0xffff), (short)(rec.w >> 16), row);
                                accumulate_se3(accR, row);
                            }
                        }
                    }
                }
            }
            else if(is_solver)
            {
                // :461-470 run even without RGB: sigma = rgbSize = 0 -> rgbError 0, count 0
                S.st.last_rgb_error = 0.f;
                S.st.last_rgb_count = 0.f;
            }
            if(level_break) break; // no arrival outstanding: every CTA takes the same branch

            // ---- reduce; the sums are digested by CTA 0 at the top of the next iteration ----
            {
                const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
                float vr = 0.f;
                if(A.rgb) vr = warp_transpose_reduce32(accR);
                s_red[warp * 64 + 32 + lane] = vr;
                __syncthreads();
                if(threadIdx.x < 64)
                {
                    float sum = 0.f;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) sum += s_red[w * 64 + threadIdx.x];
                    my_row[threadIdx.x] = sum;
                }
            }
            cta_arrive(&ctl->arrive);
            ++arr;
            pending = true;
            pending_level = lv;
        }
    }

    // ============================================================================================
    // epilogue: last solve, jump rejection (:587-591), outputs, leave the control block clean
    // ============================================================================================
    if(!pending)
    {
        // make sure every CTA has left its last wait before CTA 0 resets the counters
        cta_arrive(&ctl->arrive);
        ++arr;
    }
    if(is_solver_cta)
    {
        if(pending) solve_pending();
        else cta_wait_ge(&ctl->arrive, arr * grid);
    }
    if(is_solver)
    {
        if(A.rgb)
        {
            const float d[3] = {S.tcurr[0] - A.tprev[0], S.tcurr[1] - A.tprev[1], S.tcurr[2] - A.tprev[2]};
            if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
            {
                for(int i = 0; i < 9; i++) S.Rcurr[i] = A.Rprev[i];
                for(int i = 0; i < 3; i++) S.tcurr[i] = A.tprev[i];
            }
        }
        TrackOutput * out = A.out;
        for(int i = 0; i < 3; i++) out->trans[i] = S.tcurr[i];
        for(int i = 0; i < 9; i++) out->rot[i] = S.Rcurr[i];
        out->st = S.st;
        out->status = 1;
        __threadfence_system();
        // every other CTA has made its last arrival and waits on nothing more
        ctl->arrive = 0;
        ctl->release = 0;
        ctl->arrive_b = 0;
        for(int i = 0; i < kMaxIters; i++) { ctl->rgb_cnt[i][0] = 0; ctl->rgb_cnt[i][1] = 0; }
    }
}

struct DeviceTrack
{
    TrackCtl * ctl;
    float * partials;
    TrackOutput * out; // pinned
    int grid;
    int corr_slots;
    size_t smem_bytes;
};

} // namespace

int device_track_init(ef_tracker * t)
{
    DeviceTrack * d = new DeviceTrack();
    memset(d, 0, sizeof(*d));
    t->track_state = d;
    d->grid = t->num_sms;
    const int groups0 = (t->width / 4) * t->height;
    const int per_cta = (groups0 + d->grid - 1) / d->grid;
    d->corr_slots = (per_cta + kThreads - 1) / kThreads * 4;
    d->smem_bytes = (size_t)d->corr_slots * kThreads * sizeof(int4);
    cudaError_t e = cudaMalloc((void **)&d->ctl, sizeof(TrackCtl));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->ctl, 0, sizeof(TrackCtl), t->stream);
    if(e == cudaSuccess) e = cudaMalloc((void **)&d->partials, (size_t)d->grid * 64 * sizeof(float));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->partials, 0, (size_t)d->grid * 64 * sizeof(float), t->stream);
    if(e == cudaSuccess) e = cudaHostAlloc((void **)&d->out, sizeof(TrackOutput), cudaHostAllocMapped);
    if(e == cudaSuccess && d->smem_bytes <= 200 * 1024)
        e = cudaFuncSetAttribute(k_track, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_bytes);
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
    if(d->ctl) cudaFree(d->ctl);
    if(d->partials) cudaFree(d->partials);
    if(d->out) cudaFreeHost(d->out);
    delete d;
    t->track_state = nullptr;
}

int device_track_launch(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    if(d->smem_bytes > 200 * 1024)
    {
        t->err = "image too large for the shared-memory correspondence store of EF_SOLVE_DEVICE";
        return EF_ERR_UNSUPPORTED;
    }
    TrackArgs A;
    memset(&A, 0, sizeof(A));
    const int iterations[kNumPyrs] = {fast_odom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0}; // :384-386
    if(iterations[0] + iterations[1] + iterations[2] > kMaxIters) return EF_ERR_INVALID_ARGUMENT;
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
    }
    A.so3_last = t->last_next_image[2];
    A.so3_next = t->next_image[2];
    {
        const double K2[9] = {A.lvl[2].fx, 0, A.lvl[2].cx, 0, A.lvl[2].fy, A.lvl[2].cy, 0, 0, 1};
        double Ki[9];
        hm::inverse33(K2, Ki);
        for(int i = 0; i < 9; i++) A.so3_kinv[i] = (float)Ki[i];
    }
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
    A.ctl = d->ctl;
    A.partials = d->partials;
    A.out = d->out; // UVA: pinned + mapped host memory is addressable from the device
    A.corr_slots = d->corr_slots;

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
    if(d->out->status != 1)
    {
        t->err = "track kernel produced no result";
        return EF_ERR_BAD_STATE;
    }
    memcpy(trans, d->out->trans, sizeof(d->out->trans));
    memcpy(rot, d->out->rot, sizeof(d->out->rot));
    t->st = d->out->st;
    return EF_OK;
}

} // namespace ef
