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

constexpr int kThreads = 640;    // 20 warps = 5 per scheduler (96 registers each); 640x480 level 0 has 519 four-pixel groups per CTA -> one pass
constexpr int kWarps = kThreads / 32;
constexpr int kRedThreads = 512; // threads of CTA 0 used by the final cross-CTA sum (8 parts x 64 slots)
constexpr int kMaxIters = 32;    // SE3 iterations per call (19 in the reference schedule)
constexpr int kMaxGrid = 160;    // final-reduce unroll bound (B200: 148 SMs)
constexpr int kDbgStamps = 10;

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
    double K_inv[9];                              // host double inverse of K (RGBDOdometry.cpp:428)
};

// one 128-byte line: sector s = words [8s, 8s+7), word 8s+7 = epoch.  28 payload floats.
struct alignas(128) ParamLine
{
    unsigned w[32];
};
// SE3 payload: Rcurr[9] tcurr[3] krkinv[9] kt[3];  SO3 payload: H[9] krlr[9] done
constexpr int kPayload = 28;

struct alignas(128) TrackCtl
{
    unsigned arrive;  unsigned pad0[31];          // solve barrier: arrivals (monotonic within a launch)
    ParamLine line;                               // parameters + epoch, written by the solver thread
    unsigned long long bar_b[kMaxIters];          // per SE3 iteration: arrivals | count << 8 | sigma << 32
    double last_S[27];                            // combined normal equations of the last solve (for lastA / lastb)
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
    long long * dbg;                              // optional clock64 stamps of CTA 0 (EF_TRACK_TIMING=1)
};

// ---- memory-ordering primitives (PTX memory model, gpu scope) ----
__device__ __forceinline__ unsigned ld_relaxed(const unsigned * p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned * p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned * p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel()
{
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// all threads of the CTA call; the CTA's prior writes are ordered before the arrival (bar.sync + release)
__device__ __forceinline__ void cta_arrive(unsigned * counter)
{
    __syncthreads();
    if(threadIdx.x == 0) red_release_add(counter, 1u);
}

// CTA 0 only: wait until `target` arrivals are visible, then acquire
__device__ __forceinline__ void cta_wait_arrivals(const unsigned * counter, unsigned target)
{
    if(threadIdx.x == 0)
    {
        while(ld_relaxed(counter) < target) { }
        fence_acq_rel();
    }
    __syncthreads();
}

// solver thread: payload first, fence, then the four epoch words (each sector is validated on its own)
__device__ __forceinline__ void publish_line(ParamLine * L, const float * payload, unsigned epoch)
{
#pragma unroll
    for(int i = 0; i < kPayload; i++) st_relaxed(&L->w[(i / 7) * 8 + (i % 7)], __float_as_uint(payload[i]));
    fence_acq_rel();
#pragma unroll
    for(int s = 0; s < 4; s++) st_relaxed(&L->w[s * 8 + 7], epoch);
}

// all threads call: warp 0 spins on the line until every sector shows `epoch`, then the payload is in smem.
// A 32-byte sector is read atomically, and a sector's epoch is written after (fence) its payload, so a
// sector that shows the new epoch also shows the new payload.
__device__ __forceinline__ void wait_line(const ParamLine * L, unsigned epoch, float * s_payload)
{
    if(threadIdx.x < 32)
    {
        const unsigned lane = threadIdx.x;
        unsigned v;
        do
        {
            v = ld_relaxed(&L->w[lane]);
        } while(!__all_sync(kFullMask, ((lane & 7u) != 7u) || v == epoch));
        if((lane & 7u) != 7u) s_payload[(lane >> 3) * 7 + (lane & 7u)] = __uint_as_float(v);
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
// double-precision register needs do not inflate the per-pixel phases of the kernel.  s_final: 64 floats in
// shared memory = ICP accumulator (29, padded to 32) followed by the RGB accumulator.
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
                Sm[k] = (j == 6) ? ((double)s_final[32 + k] + w * (double)s_final[k]) : ((double)s_final[32 + k] + w * w * (double)s_final[k]);
            }
        }
    }
    else
    {
        const int o = icp ? 0 : 32;
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

// CTA 0 only, after all arrivals: add the partial rows in CTA-index order -> s_final[SLOTS].
// Loads are issued as one unrolled batch (independent), the adds stay in row order.
template<int SLOTS>
__device__ __forceinline__ void cta0_final_reduce(const float * __restrict__ partials, float * s_red, float * s_final)
{
    constexpr int kParts = kRedThreads / SLOTS;
    constexpr int kPer = (kMaxGrid + kParts - 1) / kParts;
    const int slot = threadIdx.x % SLOTS, part = threadIdx.x / SLOTS;
    if(threadIdx.x < kRedThreads)
    {
        float v[kPer];
#pragma unroll
        for(int i = 0; i < kPer; i++)
        {
            const unsigned b = part + i * kParts;
            v[i] = (b < gridDim.x) ? __ldcg(partials + b * 64 + slot) : 0.f;
        }
        float s = 0.f;
#pragma unroll
        for(int i = 0; i < kPer; i++) s += v[i];
        s_red[part * SLOTS + slot] = s;
    }
    __syncthreads();
    if(threadIdx.x < SLOTS)
    {
        float tot = 0.f;
#pragma unroll
        for(int p = 0; p < kParts; p++) tot += s_red[p * SLOTS + threadIdx.x];
        s_final[threadIdx.x] = tot;
    }
    __syncthreads();
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

__global__ void __launch_bounds__(kThreads, 1) k_track(const TrackArgs A)
{
    extern __shared__ int4 s_corr[];            // [corr_slots][kThreads]
    __shared__ float s_red[kWarps * 64];
    __shared__ float s_final[64];
    __shared__ float s_par[kPayload];
    __shared__ int s_cnt, s_sig;
    __shared__ int s_wcnt[kWarps], s_wsig[kWarps];

    TrackCtl * ctl = A.ctl;
    const unsigned grid = gridDim.x;
    const bool is_solver_cta = (blockIdx.x == 0);
    const bool is_solver = is_solver_cta && threadIdx.x == 0;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    float * my_row = A.partials + (size_t)blockIdx.x * 64;

    // barrier bookkeeping, tracked identically by every thread of the grid
    unsigned rel = 0; // parameter publications (epochs)
    unsigned arr = 0; // completed arrival rounds on ctl->arrive

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
        if(A.dbg && is_solver && dbg_it < kMaxIters) A.dbg[dbg_it * kDbgStamps + k] = clock64();
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

        const int npix = L.rows * L.cols;
        const int per_cta = (npix + grid - 1) / grid;
        const int p_begin = min(npix, (int)blockIdx.x * per_cta), p_end = min(npix, p_begin + per_cta);

        for(int it = 0; it <= 10; it++)
        {
            // ---- CTA 0: digest the previous evaluation (:348-380), publish the next homography ----
            if(is_solver_cta)
            {
                if(it > 0)
                {
                    cta_wait_arrivals(&ctl->arrive, arr * grid);
                    cta0_final_reduce<16>(A.partials, s_red, s_final);
                }
                if(is_solver)
                {
                    int done = 0;
                    if(it > 0) done = solve_so3(S, Z, s_final, it);
                    float payload[kPayload];
                    make_so3_payload(Z, payload, done, L.fx, L.fy, L.cx, L.cy, L.K_inv);
                    publish_line(&ctl->line, payload, rel + 1);
                }
            }
            ++rel;
            wait_line(&ctl->line, rel, s_par);
            const bool done = s_par[18] != 0.f;
            P.image_basis = mat_from(s_par);
            P.krlr = mat_from(s_par + 9);
            __syncthreads(); // s_par is rewritten by the next wait_line
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
            if(lane < 16) s_red[warp * 16 + lane] = lane_value;
            __syncthreads();
            if(threadIdx.x < 16)
            {
                float s = 0.f;
#pragma unroll
                for(int w = 0; w < kWarps; w++) s += s_red[w * 16 + threadIdx.x];
                my_row[threadIdx.x] = s;
            }
            cta_arrive(&ctl->arrive);
            ++arr;
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
    bool pending = false; // an arrival round whose sums CTA 0 has not digested yet
    int pending_level = 0;

    // CTA 0: wait for the outstanding arrival round, add the partial rows in CTA order, run the
    // reference's host step (:541-583) in the solver thread
    auto solve_pending = [&]() {
        cta_wait_arrivals(&ctl->arrive, arr * grid);
        stamp(1);
        cta0_final_reduce<64>(A.partials, s_red, s_final);
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
        const size_t plane = (size_t)L.rows * cols;

        float lastRGBError = FLT_MAX;      // every thread tracks it for the uniform rgb_only break (:464)
        bool first_of_level = true;

        // this CTA's contiguous run of 4-pixel groups
        const int gpr = cols >> 2;
        const int ngroups = gpr * L.rows;
        const int per_cta = (ngroups + grid - 1) / grid;
        const int g_begin = min(ngroups, (int)blockIdx.x * per_cta), g_end = min(ngroups, g_begin + per_cta);

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
                    publish_line(&ctl->line, payload, rel + 1);
                }
            }
            pending = false;
            first_of_level = false;
            ++rel;
            wait_line(&ctl->line, rel, s_par);
            IP.Rcurr = mat_from(s_par);
            IP.tcurr = make_float3(s_par[9], s_par[10], s_par[11]);
            RgbResParams RPi = RP;
            RPi.krkinv = mat_from(s_par + 12);
            RPi.kt = make_float3(s_par[21], s_par[22], s_par[23]);
            __syncthreads(); // s_par is rewritten by the next wait_line
            stamp(4);

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
                const size_t o = (size_t)y * cols + x0;

                // stage 0: every load that does not depend on arithmetic
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a, d = a, e = a, f = a;
                if(A.icp)
                {
                    a = *reinterpret_cast<const float4 *>(L.vc + o);
                    b = *reinterpret_cast<const float4 *>(L.vc + plane + o);
                    c = *reinterpret_cast<const float4 *>(L.vc + 2 * plane + o);
                    d = *reinterpret_cast<const float4 *>(L.nc + o);
                    e = *reinterpret_cast<const float4 *>(L.nc + plane + o);
                    f = *reinterpret_cast<const float4 *>(L.nc + 2 * plane + o);
                }
                // the 16-pixel border of RGBResidual (:779-783) excludes whole groups: x0 is a multiple of 4
                const bool rgb_group = A.rgb && y >= 16 && y < L.rows - 16 && x0 >= 16 && x0 < cols - 16;
                short4 gx4 = make_short4(0, 0, 0, 0), gy4 = make_short4(0, 0, 0, 0);
                float4 d14 = make_float4(0.f, 0.f, 0.f, 0.f);
                unsigned ni4 = 0;
                unsigned m0 = 0, m1 = 0, m2 = 0;
                if(rgb_group)
                {
                    gx4 = *reinterpret_cast<const short4 *>(L.dIdx + o);
                    gy4 = *reinterpret_cast<const short4 *>(L.dIdy + o);
                    d14 = *reinterpret_cast<const float4 *>(L.next_depth + o);
                    // 4 rows x 12 bytes [x0-4, x0+8) of the next image: non-zero masks, ANDed over the rows (:787-793)
                    m0 = m1 = m2 = 0xffffffffu;
#pragma unroll
                    for(int r = -2; r < 2; r++)
                    {
                        const unsigned * wp = reinterpret_cast<const unsigned *>(L.next_image + (size_t)(y + r) * cols + x0 - 4);
                        const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                        m0 &= __vcmpne4(w0, 0u);
                        m1 &= __vcmpne4(w1, 0u);
                        m2 &= __vcmpne4(w2, 0u);
                        if(r == 0) ni4 = w1;
                    }
                }

                // stage 1: projections
                float3 vg[4];
                int ux[4], uy[4];
                bool in1[4] = {false, false, false, false};
                const float vx[4] = {a.x, a.y, a.z, a.w}, vy[4] = {b.x, b.y, b.z, b.w}, vz[4] = {c.x, c.y, c.z, c.w};
                const float nx[4] = {d.x, d.y, d.z, d.w}, ny[4] = {e.x, e.y, e.z, e.w}, nz[4] = {f.x, f.y, f.z, f.w};
                if(A.icp)
                {
#pragma unroll
                    for(int k = 0; k < 4; k++) in1[k] = icp_project(IP, make_float3(vx[k], vy[k], vz[k]), vg[k], ux[k], uy[k]);
                }
                const short gxs[4] = {gx4.x, gx4.y, gx4.z, gx4.w}, gys[4] = {gy4.x, gy4.y, gy4.z, gy4.w};
                const float d1s[4] = {d14.x, d14.y, d14.z, d14.w};
                int u0[4], v0[4];
                float td1[4];
                bool in2[4];
#pragma unroll
                for(int k = 0; k < 4; k++)
                {
                    in2[k] = rgb_group && rgb_gate(RPi, x0 + k, y, gxs[k], gys[k], d1s[k]) && window_ok(m0, m1, m2, k);
                    if(in2[k]) in2[k] = rgb_warp(RPi, x0 + k, y, d1s[k], u0[k], v0[k], td1[k]);
                }

                // stage 2: gathers
                float3 vp[4], np[4];
#pragma unroll
                for(int k = 0; k < 4; k++)
                {
                    if(in1[k])
                    {
                        const size_t q = (size_t)uy[k] * cols + ux[k];
                        vp[k].x = __ldg(L.vp + q); vp[k].y = __ldg(L.vp + plane + q); vp[k].z = __ldg(L.vp + 2 * plane + q);
                        np[k].x = __ldg(L.np + q); np[k].y = __ldg(L.np + plane + q); np[k].z = __ldg(L.np + 2 * plane + q);
                    }
                }
                float d0s[4];
                unsigned ls[4];
#pragma unroll
                for(int k = 0; k < 4; k++)
                {
                    if(in2[k])
                    {
                        const size_t q = (size_t)v0[k] * cols + u0[k];
                        d0s[k] = __ldg(L.last_depth + q);
                        ls[k] = __ldg(L.last_image + q);
                    }
                }

                // stage 3: math
#pragma unroll
                for(int k = 0; k < 4; k++)
                {
                    float row[7];
                    if(in1[k] && icp_finish(IP, vg[k], make_float3(nx[k], ny[k], nz[k]), vp[k], np[k], row)) accumulate_se3(accI, row);
                }
                if(A.rgb)
                {
#pragma unroll
                    for(int k = 0; k < 4; k++)
                    {
                        const bool ok = in2[k] && rgb_accept(RPi, td1[k], d0s[k], (uint8_t)ls[k]);
                        int4 rec;
                        rec.x = -1;
                        rec.y = rec.z = rec.w = 0;
                        if(ok)
                        {
                            const float diff = static_cast<float>((ni4 >> (8 * k)) & 0xffu) - static_cast<float>(ls[k]); // reduce.cu:827
                            rec.x = (u0[k] & 0xffff) | (v0[k] << 16);
                            rec.y = __float_as_int(diff);
                            rec.z = __float_as_int(d0s[k]);
                            rec.w = ((int)gxs[k] & 0xffff) | ((int)gys[k] << 16);
                            cnt += 1;
                            sig += (int)(diff * diff); // reduce.cu:830
                        }
                        s_corr[(slot * 4 + k) * kThreads + threadIdx.x] = rec;
                    }
                }
            }
            stamp(5);

            // ICP sums leave the registers before the photometric phase needs its own 29
            {
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
                // ---- barrier B: arrivals | count << 8 | sigma << 32 in one 64-bit word.  Integer adds are exact in
                //      any order; sigma wraps mod 2^32 in the top bits exactly like the reference's int sum ----
                cnt = __reduce_add_sync(kFullMask, cnt);
                sig = __reduce_add_sync(kFullMask, sig);
                if(lane == 0) { s_wcnt[warp] = cnt; s_wsig[warp] = sig; }
                __syncthreads();
                if(threadIdx.x == 0)
                {
                    unsigned c = 0, s = 0;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) { c += (unsigned)s_wcnt[w]; s += (unsigned)s_wsig[w]; }
                    unsigned long long * word = &ctl->bar_b[cnt_slot];
                    atomicAdd(word, 1ull | ((unsigned long long)c << 8) | ((unsigned long long)s << 32));
                    unsigned long long v;
                    do
                    {
                        v = ld_relaxed64(word);
                    } while((unsigned)(v & 0xffull) < grid);
                    s_cnt = (int)((v >> 8) & 0xffffffull);
                    s_sig = (int)(unsigned)(v >> 32);
                }
                __syncthreads();
                const int rgbSize = s_cnt, sigma = s_sig;
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
                                const int pu = rec.x & 0xffff, pv = rec.x >> 16;
                                const float Z = __int_as_float(rec.z);
                                const float3 cp = project_point(pu, pv, Z, SP.inv_fx, SP.inv_fy, SP.cx, SP.cy);
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
                S.last_rgb_error = 0.f;
                S.last_rgb_count = 0.f;
            }
            if(level_break) break; // no arrival outstanding: every CTA takes the same branch
            stamp(7);

            // ---- reduce; the sums are digested by CTA 0 at the top of the next iteration ----
            {
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
            stamp(8);
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
        // make sure every CTA has left its last wait before CTA 0 resets the control block
        cta_arrive(&ctl->arrive);
        ++arr;
    }
    if(is_solver_cta)
    {
        if(pending) solve_pending();
        else cta_wait_arrivals(&ctl->arrive, arr * grid);
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
        // every other CTA has made its last arrival and waits on nothing more
        ctl->arrive = 0;
#pragma unroll
        for(int s = 0; s < 4; s++) ctl->line.w[s * 8 + 7] = 0;
        for(int i = 0; i < kMaxIters; i++) ctl->bar_b[i] = 0ull;
    }
}

struct DeviceTrack
{
    long long * dbg;      // device, kMaxIters * kDbgStamps stamps (EF_TRACK_TIMING=1)
    double dbg_acc[kMaxIters][kDbgStamps];
    long long dbg_n;
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
    d->grid = t->num_sms < kMaxGrid ? t->num_sms : kMaxGrid;
    const int groups0 = (t->width / 4) * t->height;
    const int per_cta = (groups0 + d->grid - 1) / d->grid;
    d->corr_slots = (per_cta + kThreads - 1) / kThreads * 4;
    d->smem_bytes = (size_t)d->corr_slots * kThreads * sizeof(int4);
    cudaError_t e = cudaMalloc((void **)&d->ctl, sizeof(TrackCtl));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->ctl, 0, sizeof(TrackCtl), t->stream);
    if(e == cudaSuccess) e = cudaMalloc((void **)&d->partials, (size_t)d->grid * 64 * sizeof(float));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->partials, 0, (size_t)d->grid * 64 * sizeof(float), t->stream);
    if(e == cudaSuccess) e = cudaHostAlloc((void **)&d->out, sizeof(TrackOutput), cudaHostAllocMapped);
    const char * env = getenv("EF_TRACK_TIMING");
    if(e == cudaSuccess && env && env[0] == '1')
    {
        e = cudaMalloc((void **)&d->dbg, sizeof(long long) * kMaxIters * kDbgStamps);
        if(e == cudaSuccess) e = cudaMemsetAsync(d->dbg, 0, sizeof(long long) * kMaxIters * kDbgStamps, t->stream);
    }
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
    if(d->dbg)
    {
        if(d->dbg_n > 0)
        {
            // stamps (cycles of CTA 0's SM): 0 loop top | 1 arrivals seen | 2 final reduce | 3 solve | 4 params published
            // + loaded | 5 phase A | 6 barrier B | 7 phase B | 8 reduce + arrive
            fprintf(stderr, "[ef_track timing] avg cycles per SE3 iteration over %lld calls (wait reduce solve publish | phaseA barB phaseB arrive | total)\n",
                    d->dbg_n);
            for(int it = 0; it < kMaxIters; it++)
            {
                const double * a = d->dbg_acc[it];
                if(a[4] == 0) continue;
                const double n = (double)d->dbg_n;
                const bool first = a[1] == 0;
                fprintf(stderr, "  it %2d: %7.0f %7.0f %7.0f %7.0f | %7.0f %7.0f %7.0f %7.0f | %8.0f\n", it, first ? 0.0 : (a[1] - a[0]) / n,
                        first ? 0.0 : (a[2] - a[1]) / n, first ? 0.0 : (a[3] - a[2]) / n, (a[4] - (first ? a[0] : a[3])) / n, (a[5] - a[4]) / n,
                        (a[6] - a[5]) / n, (a[7] - a[6]) / n, (a[8] - a[7]) / n, (a[8] - a[0]) / n);
            }
        }
        cudaFree(d->dbg);
    }
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
    if(d->smem_bytes > 200 * 1024 || (size_t)t->width * t->height >= (1u << 24))
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
    A.ctl = d->ctl;
    A.partials = d->partials;
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
        long long h[kMaxIters * kDbgStamps];
        if(cudaMemcpy(h, d->dbg, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess)
        {
            for(int it = 0; it < kMaxIters; it++)
                for(int k = 0; k < kDbgStamps; k++) d->dbg_acc[it][k] += (double)h[it * kDbgStamps + k];
            d->dbg_n++;
            cudaMemset(d->dbg, 0, sizeof(h));
        }
    }
    return EF_OK;
}

} // namespace ef
