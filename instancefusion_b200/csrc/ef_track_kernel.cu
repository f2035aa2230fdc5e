// ef_track_kernel.cu -- EF_SOLVE_DEVICE: the whole of RGBDOdometry::getIncrementalTransformation
// (elasticfusionpublic/Core/src/Utils/RGBDOdometry.cpp:292-591) as ONE persistent cooperative kernel.
//
// Why: at 640x480 a Gauss-Newton iteration touches ~19 MB that already sits in the 126 MB L2, i.e. a
// couple of microseconds of memory time, while the reference pays 3-4 launches, 3-4 device syncs and
// blocking D2H copies per iteration (57+ host round trips per frame).  Here one CTA per SM stays
// resident for the whole solve and an iteration costs a handful of L2 round trips:
//
//   level start  every worker CTA (147 of them; pixels are dealt round-robin in 32-pixel chunks) derives dIdx / dIdy of its
//             pixels (computeDerivativeImages fused in), evaluates the ITERATION-INVARIANT photometric gates once (16-pixel
//             border, gradient magnitude, depth validity, 4x4 validity window: reduce.cu:779-811), compacts the survivors in
//             a fixed order into a candidate list in SHARED memory (pixel, intensity, gradients, depth = 12 B each) and stages
//             the current-frame vertices / normals of its pixels there too.  This overlaps the previous level's last solve.
//   phase A1  photometric association of the candidates (warp + gather + accept, reduce.cu:813-830) -> 8-byte match
//             records in shared memory; the CTA's {count, sum diff^2} leaves as ONE flagged chunk ("barrier B": the robust
//             weight needs the global sums); CTA 0 adds the 147 arrivals and answers every worker with one chunk.
//   phase A2  ICP association + 29 sums in registers, 3 pixels of a thread interleaved, no control flow; hides barrier B.
//   phase B   photometric rows from the records in shared memory -> 29 more sums.
//   reduce    transpose-reduce butterfly per warp (skipped by warps that had no pixel) -> per-CTA 58-float row ->
//             published as flagged 16-byte chunks.
//   solve     CTA 0 polls all rows with up to 12 independent loads in flight per thread, adds them in worker order
//             (deterministic); its WARP 0 runs the reference's host step in double (LDL^T, exp map, SE(3) update, pose
//             composition), every lane the whole scalar routine on broadcast data -- no shuffles --, warp 1 derives the
//             photometric warp K R K^-1, K t meanwhile; both publish their half of the parameter line as flagged chunks.
//
// No DataTerm image, no point cloud, no reduceSum launch, no host involvement until the final pose is
// stored straight into pinned host memory.  Co-residency of the spinning CTAs is guaranteed by a
// cooperative launch with gridDim <= number of SMs.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ef_hostmath.h"
#include "ef_image_px.cuh"
#include "ef_kernels.h"
#include "ef_pixel.cuh"
#include "ef_reduce.cuh"
#include "ef_tracker.h"

// This file is compiled FOUR times (csrc/Makefile).  Single launches: 256 threads per CTA (8 warps, <= 255 registers) is the
// best shape up to ~640x480, 384 threads (12 warps, <= 168 registers) from ~1280x720 on, where a thread owns ~25 pixels per
// level-0 iteration and the extra warps hide more latency than the longer reductions cost
// (profiles/r01_k_track_thread_sweep.txt); ef_track_dispatch.cu picks the variant per handle from the image size.  k sequences
// per launch: the alternating build (EF_TRACK_ALT, k_track_alt below, the default of ef_track_frames_to_model_batch) and the
// thread-group build (EF_TRACK_GROUPS = 2).
#ifndef EF_TRACK_THREADS
#define EF_TRACK_THREADS 256
#endif
// EF_TRACK_GROUPS > 1: the BATCHED build.  A CTA then holds EF_TRACK_GROUPS independent thread groups of EF_TRACK_THREADS
// threads, group g of every CTA working for sequence g of the launch (its own tracker handle: own pyramids, control block,
// shared-memory region, named barrier).  Nothing else changes -- group g of CTA 0 gathers and solves for sequence g, group
// g of the other CTAs does its pixels -- but the SM's warp schedulers now interleave the sequences by themselves: while
// the warps of one group wait for parameters (37-60 % of an iteration, profiles/r01_k_track_phase_trace_final.txt) the
// other group's warps own the issue slots.  Same instructions per sequence, so the same bits as the single launch.
#ifndef EF_TRACK_GROUPS
#define EF_TRACK_GROUPS 1
#endif
#define EF_TRACK_CAT2(a, b) a##_t##b
#define EF_TRACK_CAT(a, b) EF_TRACK_CAT2(a, b)
// EF_TRACK_ALT: the ALTERNATING build -- k sequences per launch by time-slicing the worker CTAs instead of splitting their
// threads.  Sequence s has its own solver CTA (block s); every worker CTA (all kThreads threads, one register file) runs one
// Gauss-Newton iteration of sequence 0, publishes its partial sums, runs an iteration of sequence 1, ... and by the time it is
// back at sequence 0 that sequence's solver has gathered, solved and published the next pose: the serial chain of one sequence
// (37-60 % of an iteration) is hidden behind the pixels of the others.  See k_track_alt.
#if defined(EF_TRACK_ALT)
#if EF_TRACK_GROUPS != 1
#error "EF_TRACK_ALT is a build of the single-group kernel"
#endif
#define EF_TRACK_FN(name) name##_alt
#elif EF_TRACK_GROUPS > 1
#define EF_TRACK_CATG2(a, c) a##_g##c
#define EF_TRACK_CATG(a, c) EF_TRACK_CATG2(a, c)
#define EF_TRACK_FN(name) EF_TRACK_CATG(name, EF_TRACK_GROUPS) // (one batched build per library, whatever its group size)
#else
#ifndef EF_TRACK_NAME_THREADS // the slot of ef_track_dispatch.cu this build fills (256 | 384); sweeps put other shapes into a slot
#define EF_TRACK_NAME_THREADS EF_TRACK_THREADS
#endif
#define EF_TRACK_FN(name) EF_TRACK_CAT(name, EF_TRACK_NAME_THREADS)
#endif

namespace ef
{

void EF_TRACK_FN(device_track_destroy)(ef_tracker * t);
int EF_TRACK_FN(device_track_trace)(ef_tracker * t, double * out32, long long * calls);
int EF_TRACK_FN(device_track_configure)(ef_tracker * t, int grid);
bool EF_TRACK_FN(device_track_supported)(const ef_tracker * t);

namespace
{

constexpr int kThreads = EF_TRACK_THREADS; // threads of a group; 8 warps, up to 255 registers each: measured best of {128..640} (tools/sweep.sh)
constexpr int kGroups = EF_TRACK_GROUPS;   // independent sequences per launch (thread groups per CTA)
static_assert(kGroups >= 1 && kGroups <= 2 && kThreads * kGroups <= 1024, "EF_TRACK_GROUPS");

// thread index within its group / group index (warp-uniform: kThreads is a multiple of 32)
__device__ __forceinline__ unsigned gtid() { return kGroups > 1 ? threadIdx.x % kThreads : threadIdx.x; }
__device__ __forceinline__ unsigned ggrp() { return kGroups > 1 ? threadIdx.x / kThreads : 0u; }
// barrier of one group: the whole CTA in the single build, named barrier 1 + g of kThreads threads in the batched one
// (immediate barrier numbers: ptxas then knows which barriers the kernel uses, and compute-sanitizer's synccheck can follow them)
__device__ __forceinline__ void group_sync()
{
    if constexpr(kGroups == 1) __syncthreads();
    else
    {
        static_assert(kGroups <= 2, "one immediate barrier number per group");
        if(ggrp() == 0) asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
    }
}
#ifndef EF_TRACK_ICP_BATCH
#if EF_TRACK_THREADS >= 384
#define EF_TRACK_ICP_BATCH 2 // 168 registers: two pixels in flight per thread
#else
#define EF_TRACK_ICP_BATCH 3
#endif
#endif
constexpr int kIcpBatch = EF_TRACK_ICP_BATCH; // pixels of one thread whose loads and gathers are in flight together
#ifndef EF_TRACK_RGB_BATCH
#define EF_TRACK_RGB_BATCH 3
#endif
constexpr int kRgbBatch = EF_TRACK_RGB_BATCH; // photometric candidates of one thread in flight together (1 .. 3)
#ifndef EF_TRACK_STAGE_BATCH
#define EF_TRACK_STAGE_BATCH 5
#endif
#ifndef EF_TRACK_ICP_SPLIT
#define EF_TRACK_ICP_SPLIT 0
#endif
constexpr bool kIcpSplit = EF_TRACK_ICP_SPLIT != 0;
constexpr int kWarps = kThreads / 32;
// Accumulator sets: the float sums of a launch are added in the order of EF_TRACK_SETS x kThreads VIRTUAL threads
// (EF_TRACK_SETS x kWarps virtual warps), each physical thread keeping one accumulator set per virtual thread it stands for
// (virtual thread = thread + set * kThreads; its pixels are the physical passes / candidate rounds with index = set modulo
// EF_TRACK_SETS).  The single-launch builds use one set; the batched build runs groups of 128 threads with two, which makes
// its summation tree the one of the 256-thread single launch: per handle it returns the same bits.
#ifndef EF_TRACK_SETS
#define EF_TRACK_SETS 1
#endif
constexpr int kSets = EF_TRACK_SETS;
constexpr int kVWarps = kWarps * kSets;
static_assert(kSets == 1 || kSets == 2, "EF_TRACK_SETS");
constexpr int kMaxIters = 32;    // SE3 iterations per call (19 in the reference schedule)
constexpr int kDbgStamps = 13; // 0..9 per iteration; 10..12 level start (first iteration of a level only)
constexpr int kRowChunks = 20;   // a partial row = 20 x (3 floats + flag) = 60 floats >= 29 ICP + 29 RGB sums
constexpr int kRowFloats = kRowChunks * 3;
constexpr int kSo3Chunks = 4;    // SO3 rows carry 11 floats
constexpr int kLineChunks = 8;   // the parameter line = 8 x (3 floats + flag) = 24 floats
constexpr int kPayload = kLineChunks * 3;
constexpr int kParts = kThreads / 64;  // final cross-CTA sum: kVParts x 64 slots, a thread adding kSets of them
constexpr int kVParts = kParts * kSets;
constexpr int kGatherBatch = 12; // row chunks a CTA-0 thread requests before it examines the first (147 x 20 / 256 = 11.5)
#ifndef EF_TRACK_COMPACT_GROUP
#define EF_TRACK_COMPACT_GROUP 3
#endif
constexpr int kCompactGroup = EF_TRACK_COMPACT_GROUP; // passes of the level-start compaction evaluated together
constexpr int kMaxGrid = 255;    // CTAs of one launch (sizes the control buffers)
constexpr unsigned kNoMatch = 0xffffffffu;
constexpr int kCandBytes = 20;   // 12-byte candidate + 8-byte match record
constexpr int kIcpBytes = 24;    // current-frame vertex + normal of a pixel (icp_in_smem)
constexpr int kMaxDynSmem = 200 * 1024;

static_assert(kThreads % 32 == 0 && kThreads >= 128 && kThreads <= 1024, "EF_TRACK_THREADS");
static_assert(kMaxGrid == 255 && kRowChunks == 20, "kSlotCopy / kRowsChunks");

struct LevelArgs
{
    const float * vc, * nc, * vp, * np;           // 3-plane maps, dense
    const float * last_depth, * next_depth;
    const uint8_t * last_image, * next_image;
    int16_t * dIdx, * dIdy;                       // written here when TrackArgs::make_derivatives (computeDerivativeImages fused in)
    int rows, cols;
    float fx, fy, cx, cy;                         // level intrinsics
    float inv_fx, inv_fy;                         // host 1.0f / f (cudafuncs.cu:671)
    float min_scale;
    int iterations;
    double K_inv[9];                              // host double inverse of K (RGBDOdometry.cpp:428)
};

// Flagged 16-byte chunks (the NCCL "LL" idea): a chunk is written with ONE 128-bit store and read with ONE
// 128-bit load, so its three payload words and its flag are always observed together.  No fence, no separate
// "ready" counter: whoever polls a chunk gets the data with the same load that tells it the data is there.
// The accesses are the SCALAR 128-bit forms  ld/st.relaxed.gpu.global.b128  (PTX ISA 8.3+, sm_70+): one memory
// operation of a 128-bit scalar type under the PTX memory model, not the .v4.u32 vector forms, which the model
// treats as four independent 32-bit accesses in unspecified order (round-1 review).  Both compile to the same
// LDG.E.128.STRONG.GPU / STG.E.128.STRONG.GPU; tools/chunk_torture.cu hammers the pattern (every reader checks every
// observed chunk for tearing, all SM pairs, under concurrent HBM traffic), log in profiles/r02_chunk_torture.txt.
// Flags are launch-unique epochs (launch_seq << 8 | n), so nothing has to be reset between launches.
// Nobody polls a line together with more than ~5 other CTAs (148 CTAs spinning on one line serialise in its L2 slice):
//   par      the parameter line (8 chunks), kReplicas copies 256 bytes apart; worker w reads copy w % kReplicas.
//            Chunks 0..3 = pose (Rcurr, tcurr), published as soon as the solve is done; chunks 4..7 = photometric warp
//            (K R K^-1, K t), published later -- the workers start their ICP pixels in between.
//   bslot[w] worker w's barrier-B arrival {count, sum diff^2}, polled by thread w of CTA 0
//   bres[w]  the barrier-B result for worker w {sigma, rgbError, count}, written by thread w of CTA 0
//   rows     worker w's partial sums of a round at rows[w * chunks ..), polled by CTA 0
#ifndef EF_TRACK_REPLICAS
#define EF_TRACK_REPLICAS 32 // measured 4 .. 148: k_track 0.1710 ms with 8, 0.1695 with 16, 0.1681 with 32 / 64, 0.1735 with one line per CTA
#endif
constexpr int kReplicas = EF_TRACK_REPLICAS;
constexpr int kReplicaStride = 16; // chunks
// The parameter line exists twice and publication `epoch` goes to copy epoch & 1.  Publication e + 2 needs an acknowledgement of
// e + 1 from every worker (its rows or its barrier-B arrival), which a worker sends after it consumed e, so a copy is never
// overwritten before everybody has read it -- however late a worker gets to its wait (the alternating build's workers arrive
// long after the publication).  With ONE copy the pose of the first Gauss-Newton iteration followed the "SO(3) is over" message
// with nothing in between and a worker that was not already polling would have missed the latter.
constexpr int kParCopyStride = kReplicas * kReplicaStride; // chunks between the two copies
// control block of a handle: 2 parameter copies | 2 x kSlotCopy barrier-B arrivals (two copies by iteration parity in the symmetric
// body) | kSlotCopy barrier-B answers; rows: two copies of kMaxGrid rows (by round parity, symmetric body)
constexpr size_t kSlotCopy = (255 + 7) & ~7;
constexpr size_t kCtlChunks = 2 * (size_t)kParCopyStride + 3 * kSlotCopy;
constexpr size_t kRowsChunks = 2 * (size_t)255 * 20;
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
    int cand_cap;                                 // capacity of the shared-memory candidate list (largest level)
    int lvl_cap[kNumPyrs];                        // pixels a worker CTA owns at each level (capacity of that level's lists)
    int lvl_off[kNumPyrs];                        // byte offset of each level's lists in dynamic shared memory
    int prework;                                  // all levels' lists fit side by side: finer levels are prepared during the waits of coarser ones
    int make_derivatives;                         // dIdx / dIdy are not valid yet: compute them at level start
    int icp_in_smem;                              // the CTA's current-frame vertices / normals fit shared memory beside the candidates
    int group_smem_bytes;                         // dynamic shared memory of one thread group (multiple of 16)
    int cand_base;                                // symmetric body: byte offset of the candidate lists (behind the gathered rows)
    float prev_icp_error, prev_icp_count, prev_so3_error, prev_so3_count, prev_rgb_error, prev_rgb_count;
    unsigned epoch_base;                          // launch_seq << 8
    uint4 * par;                                  // kReplicas parameter lines
    uint4 * bslot, * bres;                        // one flagged chunk per worker each
    uint4 * rows;                                 // flagged chunks: worker w's row of a round at rows[w * chunks ..)
    TrackOutput * out;
    long long * dbg;                              // optional clock64 stamps of every CTA (EF_TRACK_TIMING=1)
};

// ---- 128-bit relaxed (L2-coherent) SCALAR accesses: one memory operation each (see above) ----
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 * p)
{
    uint4 v;
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%4];\n\tmov.b128 {%0, %1, %2, %3}, q;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint4 * p, const uint4 & v)
{
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// warp 0 of CTA 0, payload already in shared memory: chunks [first, first + n) of every replica of the parameter
// line; with n = 4 that is ONE store per lane
__device__ __forceinline__ void warp_publish(uint4 * par, const float * s_payload, int first, int n, unsigned epoch)
{
    __syncwarp();
    par += (epoch & 1u) * kParCopyStride;
    for(int i = gtid() & 31; i < kReplicas * n; i += 32)
    {
        const int r = i / n, c = first + (i - r * n);
        st_relaxed_v4(par + r * kReplicaStride + c,
                      make_uint4(__float_as_uint(s_payload[3 * c]), __float_as_uint(s_payload[3 * c + 1]), __float_as_uint(s_payload[3 * c + 2]), epoch));
    }
}

// all threads of a worker CTA call: lanes 0..n-1 of warp 0 spin on chunks [first, first + n) of this CTA's replica until
// they show `epoch`; payload -> smem
__device__ __forceinline__ void wait_chunks(const uint4 * line, int first, int n, unsigned epoch, float * s_payload)
{
    if(gtid() < 32)
    {
        const int lane = gtid();
        line += (epoch & 1u) * kParCopyStride;
        uint4 v = make_uint4(0, 0, 0, epoch);
        do
        {
            if(lane < n) v = ld_relaxed_v4(line + first + lane);
        } while(!__all_sync(kFullMask, v.w == epoch));
        if(lane < n)
        {
            float * d = s_payload + 3 * (first + lane);
            d[0] = __uint_as_float(v.x);
            d[1] = __uint_as_float(v.y);
            d[2] = __uint_as_float(v.z);
        }
    }
    group_sync();
}

// worker CTA: publish this CTA's partial sums (already in shared memory) as flagged chunks
__device__ __forceinline__ void publish_row(uint4 * my_row, const float * s_row, int chunks, unsigned epoch)
{
    if((int)gtid() < chunks)
    {
        const int c = gtid();
        st_relaxed_v4(my_row + c, make_uint4(__float_as_uint(s_row[3 * c]), __float_as_uint(s_row[3 * c + 1]), __float_as_uint(s_row[3 * c + 2]), epoch));
    }
}

// CTA 0: collect every worker's row -- all of a thread's chunks are requested before the first one is examined, so
// the round costs one L2 round trip after the last row landed -- then add the rows in worker order -> s_final[64]
__device__ __forceinline__ void gather_rows(const uint4 * rows, int workers, int chunks, unsigned epoch, float * s_rows /*[workers][chunks * 3]*/,
                                            float * s_red, float * s_final)
{
    const int total = workers * chunks;
    for(int i0 = gtid(); i0 < total; i0 += kGatherBatch * kThreads)
    {
        uint4 v[kGatherBatch];
        unsigned todo = 0;
#pragma unroll
        for(int k = 0; k < kGatherBatch; k++)
            if(i0 + k * kThreads < total) todo |= 1u << k;
        // rounds: every chunk still missing is requested again, all requests of a round in flight together
        while(todo)
        {
#pragma unroll
            for(int k = 0; k < kGatherBatch; k++)
                if(todo & (1u << k)) v[k] = ld_relaxed_v4(rows + i0 + k * kThreads);
#pragma unroll
            for(int k = 0; k < kGatherBatch; k++)
                if((todo & (1u << k)) && v[k].w == epoch)
                {
                    todo &= ~(1u << k);
                    float * d = s_rows + 3 * (i0 + k * kThreads);
                    d[0] = __uint_as_float(v[k].x);
                    d[1] = __uint_as_float(v[k].y);
                    d[2] = __uint_as_float(v[k].z);
                }
        }
    }
    group_sync();
    const int nfl = chunks * 3;
    const int slot = gtid() & 63, part = gtid() >> 6;
    if(part < kParts)
    {
#pragma unroll
        for(int q = 0; q < kSets; q++)
        {
            const int vpart = part + q * kParts;
            float s = 0.f;
            if(slot < nfl)
                for(int w = vpart; w < workers; w += kVParts) s += s_rows[w * nfl + slot];
            s_red[vpart * 64 + slot] = s;
        }
    }
    group_sync();
    if(gtid() < 64)
    {
        float tot = 0.f;
#pragma unroll
        for(int p = 0; p < kVParts; p++) tot += s_red[p * 64 + gtid()];
        s_final[gtid()] = tot;
    }
    group_sync();
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
// solver state: shared memory of CTA 0, read by the lanes of its warps 0 and 1 (broadcast) and written by lane 0
// ------------------------------------------------------------------------------------------------
struct Solver
{
    double resultRt[32];                          // 3x4 block row-major in [0, 12); [12, 32) is scratch so that all lanes may store
    double last_S[28];                            // combined normal equations of the last solve (lastA / lastb); [27] is scratch
    double so3_R[9], so3_lastR[9];
    float Rcurr[9], tcurr[3];
    float Rprev[9], tprev[3];                     // copies of the launch constants (indexed per lane by the warp solver)
    float so3_R_lr[9], so3_lastError, so3_lastCount;
    float last_icp_error, last_icp_count, last_rgb_error, last_rgb_count, last_so3_error, last_so3_count;
    int so3_iterations, se3_iterations[3];
};

// 1 / d to double precision without the IEEE-division slow path: hardware seed (2^-20) + two Newton steps
// (error < 2^-52 relative; the pivots of the SPD normal equations are normal numbers); 0 -> 0 like the scalar routine
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return (d != 0.0) ? r : 0.0;
}

// The ill-conditioned case of the 6x6 solve (RGBDOdometry.cpp:552-564): the same symmetric-pivoted LDL^T as the host-solve
// mode and the reference's Eigen ldlt() (pivot = largest remaining |diagonal|, hm::ldlt_solve).  Taken only when the
// unpivoted fast path met a pivot below 1e-10 of the largest diagonal (plane-only ICP, too few photometric matches, a
// nearly empty frame), where the order of elimination decides what comes out of the null space.  Rare, so out of line.
__device__ __noinline__ void solve_pivoted(const double * S28, double * x6)
{
    double A[36], b[6];
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
#pragma unroll
        for(int j = i; j < 7; j++)
        {
            const double v = S28[hm::acc_index(i, j)];
            if(j == 6) b[i] = v;
            else A[j * 6 + i] = A[i * 6 + j] = v;
        }
    }
    hm::ldlt_solve<double, 6>(A, b, x6);
}

// warps 0 and 1 of CTA 0: hand-over of resultRt from the solver warp to the warp that derives the photometric warp
// (out of line: both warps then wait at ONE barrier instruction -- bar.sync pairs warps by barrier number, not by address, but
//  compute-sanitizer's synccheck reports arrivals from different addresses as divergence; a call costs ~20 cycles per iteration)
__device__ __noinline__ void solver_pair_sync()
{
    if constexpr(kGroups == 1) asm volatile("bar.sync 1, 64;" ::: "memory");
    else
    {
        if(ggrp() == 0) asm volatile("bar.sync 3, 64;" ::: "memory");
        else asm volatile("bar.sync 4, 64;" ::: "memory");
    }
}


// OdometryProvider.h:35-71 rodrigues(r) for the small rotations of a tracked frame; t2 = |r|^2
__device__ __forceinline__ void rodrigues_small(double rx, double ry, double rz, double t2, double * R)
{
    if(t2 < 0.0625)
    {
        // R = cos(t) I + (1 - cos t) r^ r^T + sin(t) [r^]x  with r^ = r / t  (:52-68)
        //   = (1 - B t^2) I + B r r^T + A [r]x,  A = sin(t)/t, B = (1 - cos t)/t^2: two short alternating series in t^2
        // (|t| < 0.25 rad: 9 terms reach 2^-60), no square root, no division, no argument reduction -- a Gauss-Newton
        // update of a tracked frame is a few milliradians.  Equal to the closed form to double rounding; t -> 0 gives the
        // identity, which is also what the reference returns below DBL_EPSILON (:45).
        double A = 1.0 / 121645100408832000.0, B = 1.0 / 2432902008176640000.0; // 1/19!, 1/20!
        A = fma(-t2, A, 1.0 / 355687428096000.0);    B = fma(-t2, B, 1.0 / 6402373705728000.0);
        A = fma(-t2, A, 1.0 / 1307674368000.0);      B = fma(-t2, B, 1.0 / 20922789888000.0);
        A = fma(-t2, A, 1.0 / 6227020800.0);         B = fma(-t2, B, 1.0 / 87178291200.0);
        A = fma(-t2, A, 1.0 / 39916800.0);           B = fma(-t2, B, 1.0 / 479001600.0);
        A = fma(-t2, A, 1.0 / 362880.0);             B = fma(-t2, B, 1.0 / 3628800.0);
        A = fma(-t2, A, 1.0 / 5040.0);               B = fma(-t2, B, 1.0 / 40320.0);
        A = fma(-t2, A, 1.0 / 120.0);                B = fma(-t2, B, 1.0 / 720.0);
        A = fma(-t2, A, 1.0 / 6.0);                  B = fma(-t2, B, 1.0 / 24.0);
        A = fma(-t2, A, 1.0);                        B = fma(-t2, B, 0.5);
        const double d = fma(-B, t2, 1.0);
        const double bx = B * rx, by = B * ry, bz = B * rz;
        R[0] = fma(bx, rx, d);             R[1] = fma(bx, ry, -(A * rz));    R[2] = fma(bx, rz, A * ry);
        R[3] = fma(by, rx, A * rz);        R[4] = fma(by, ry, d);            R[5] = fma(by, rz, -(A * rx));
        R[6] = fma(bz, rx, -(A * ry));     R[7] = fma(bz, ry, A * rx);       R[8] = fma(bz, rz, d);
    }
    else
    {
        const double xr[3] = {rx, ry, rz};
        hm::rodrigues(xr, R);
    }
}

// RGBDOdometry.cpp:515-516, :541-583 -- the reference's host step after a Gauss-Newton evaluation, run by WARP 0 of CTA 0
// with NO lane parallelism: every lane runs the scalar routine on identical data (shared-memory broadcast loads), the 27
// normal-equation entries in registers.  A double shuffle costs 26 cycles and a DFMA 8 (tools/op_latency), so a version that
// spreads the entries over the lanes spends most of each LDL^T pivot moving operands (measured: 196 cycles per pivot, 3 200 per
// solve); here a pivot is reciprocal -> multiply -> fused multiply-add (~80 cycles), its ~15 trailing updates are independent
// DFMAs, and a warp-wide DFMA issues as fast as a one-lane one (1 870 cycles per solve).  No shuffles and no lane-dependent
// branches: nvcc does not re-converge `if(lane < n)` regions here and every later warp-collective would take the
// WARPSYNC.COLLECTIVE slow path (~150 cycles each, measured).  All in double like the reference.
// s_final (shared memory): ICP accumulator [0, 29) followed by the RGB accumulator [29, 58).
__device__ __forceinline__ void warp_solve_se3(Solver * S, const float * s_final, int icp, int rgb, float icp_weight, int level,
                                                      bool hand_over, long long * dbg = nullptr)
{
    long long tk[5] = {0, 0, 0, 0, 0};
    if(dbg) tk[0] = clock64();
    const int lane = gtid() & 31;
    __syncwarp();
    // combined normal equations (:547-553): lane k converts and combines entry k (one conversion per lane instead of 54 per
    // lane), parks it in last_S -- which is lastA / lastb of this solve anyway (RGBDOdometry.h:72-73) -- and every lane
    // then reads all 27 entries back as 14 broadcast 128-bit loads
    const double w = icp_weight, ww = w * w;
    const int sl = lane < 27 ? lane : 27;
    const int lr = (sl >= 7) + (sl >= 13) + (sl >= 18) + (sl >= 22) + (sl >= 25);
    const bool rhs = (sl - (lr * 7 - (lr * (lr - 1)) / 2) + lr) == 6;
    const double l_icp = (double)s_final[sl], l_rgb = (double)s_final[29 + sl];
    S->last_S[sl] = (icp && rgb) ? (rhs ? (l_rgb + w * l_icp) : (l_rgb + ww * l_icp)) : (icp ? l_icp : l_rgb);
    __syncwarp();
    double a[28];
#pragma unroll
    for(int k = 0; k < 14; k++)
    {
        const double2 v = reinterpret_cast<const double2 *>(S->last_S)[k];
        a[2 * k] = v.x;
        a[2 * k + 1] = v.y;
    }

    // x = A^-1 b (:552-564): hm::ldlt_solve_spd6_acc with the Newton reciprocal.  Beside the critical path: the largest
    // diagonal entry and the smallest pivot, which decide whether this unpivoted elimination was good enough
    double inv[6], x[6];
    double amax = 0.0, dmin = DBL_MAX;
#pragma unroll
    for(int j = 0; j < 6; j++) amax = fmax(amax, fabs(a[hm::acc_index(j, j)]));
#pragma unroll
    for(int j = 0; j < 6; j++)
    {
        dmin = fmin(dmin, a[hm::acc_index(j, j)]); // (a negative or NaN pivot also trips the test below)
        inv[j] = fast_rcp(a[hm::acc_index(j, j)]);
        double l[6];
#pragma unroll
        for(int i = j + 1; i < 6; i++) l[i] = a[hm::acc_index(j, i)] * inv[j];
#pragma unroll
        for(int i = j + 1; i < 6; i++)
        {
#pragma unroll
            for(int k = i; k < 7; k++) a[hm::acc_index(i, k)] = fma(-l[i], a[hm::acc_index(j, k)], a[hm::acc_index(i, k)]);
        }
#pragma unroll
        for(int i = j + 1; i < 6; i++) a[hm::acc_index(j, i)] = l[i];
    }
    if(dbg) tk[1] = clock64();
    // back substitution, column oriented: x(i) is final once the rows below it were applied
    double wv[6];
#pragma unroll
    for(int i = 0; i < 6; i++) wv[i] = a[hm::acc_index(i, 6)] * inv[i];
#pragma unroll
    for(int i = 5; i >= 0; i--)
    {
        x[i] = wv[i];
#pragma unroll
        for(int r = 0; r < i; r++) wv[r] = fma(-a[hm::acc_index(r, i)], x[i], wv[r]);
    }
    if(!(dmin > 1e-10 * amax)) solve_pivoted(S->last_S, x); // warp-uniform: every lane holds the same numbers
    if(dbg) tk[2] = clock64();

    // OdometryProvider.h:35-71 rodrigues(x[3:6])
    const double rx = x[3], ry = x[4], rz = x[5];
    const double t2 = rx * rx + ry * ry + rz * rz;
    double R[9];
    rodrigues_small(rx, ry, rz, t2, R);
    if(dbg) tk[3] = clock64();
    // OdometryProvider.h:73-93: resultRt = [R | x[0:3]] * resultRt
    double rt[12], N[12];
#pragma unroll
    for(int i = 0; i < 12; i++) rt[i] = S->resultRt[i];
#pragma unroll
    for(int r = 0; r < 3; r++)
#pragma unroll
        for(int c = 0; c < 4; c++)
        {
            double v = R[r * 3] * rt[c];
            v = fma(R[r * 3 + 1], rt[4 + c], v);
            v = fma(R[r * 3 + 2], rt[8 + c], v);
            N[r * 4 + c] = (c == 3) ? v + x[r] : v;
        }
    // every lane stores the same values (one transaction per entry); warp 1 derives the photometric warp from them
    __syncwarp(); // all lanes have read the old matrix (they run in lock step anyway; this states it for racecheck)
#pragma unroll
    for(int i = 0; i < 12; i++) S->resultRt[i] = N[i];
    if(hand_over) solver_pair_sync();

    // :571-583: [Rcurr | tcurr] = [Rprev | tprev] * (float(resultRt))^-1 with the Isometry3f inverse (R^T, -R^T t), float
    float Rc[9], tc[3];
    hm::compose_pose_affine12(N, S->Rprev, S->tprev, Rc, tc);
    if(dbg) tk[4] = clock64();
    if(dbg && lane == 0)
    {
        dbg[8] = tk[1] - tk[0]; // combine + LDL^T
        dbg[9] = tk[2] - tk[1]; // back substitution
        dbg[5] = tk[3] - tk[2]; // exponential map
        dbg[4] = tk[4] - tk[3]; // update + compose
    }
    if(lane == 0)
    {
#pragma unroll
        for(int i = 0; i < 9; i++) S->Rcurr[i] = Rc[i];
#pragma unroll
        for(int i = 0; i < 3; i++) S->tcurr[i] = tc[i];
        S->se3_iterations[level]++;
        if(icp)
        {
            S->last_icp_error = sqrtf(s_final[27]) / s_final[28]; // :515-516
            S->last_icp_count = s_final[28];
        }
    }
    __syncwarp();
}

// :480-481 -- pose of the next SE3 iteration (warp 0 of CTA 0) -> payload[0, 12) in shared memory
__device__ __forceinline__ void warp_make_pose(const Solver * S, float * payload /*shared*/)
{
    const int lane = gtid() & 31;
    __syncwarp();
    if(lane < 12) payload[lane] = (lane < 9) ? S->Rcurr[lane] : S->tcurr[lane - 9];
}

// :424-434 -- photometric warp of the next SE3 iteration (warp 1 of CTA 0) -> payload[12, 24) in shared memory:
// krkinv[9] kt[3] with Rt = resultRt^-1 (general affine inverse: adjugate / determinant), krkinv = K R K^-1, kt = K t in
// double (hm::rgb_warp_params_sparse is the scalar statement); every lane evaluates it (see warp_solve_se3)
__device__ __forceinline__ void warp_make_rgb_params(const Solver * S, float * payload /*shared*/, float fxf, float fyf, float cxf, float cyf,
                                                            const double * K_inv)
{
    const int lane = gtid() & 31;
    __syncwarp();
    double m[12];
#pragma unroll
    for(int i = 0; i < 12; i++) m[i] = S->resultRt[i];
    // Rt = resultRt^-1 (general affine inverse: adjugate / determinant)
    const double c0 = m[5] * m[10] - m[6] * m[9], c1 = m[2] * m[9] - m[1] * m[10], c2 = m[1] * m[6] - m[2] * m[5];
    const double c3 = m[6] * m[8] - m[4] * m[10], c4 = m[0] * m[10] - m[2] * m[8], c5 = m[2] * m[4] - m[0] * m[6];
    const double c6 = m[4] * m[9] - m[5] * m[8], c7 = m[1] * m[8] - m[0] * m[9], c8 = m[0] * m[5] - m[1] * m[4];
    const double det = m[0] * c0 + m[1] * c3 + m[2] * c6;
    const double id = 1.0 / det;
    const double Ai[9] = {c0 * id, c1 * id, c2 * id, c3 * id, c4 * id, c5 * id, c6 * id, c7 * id, c8 * id};
    double ti[3];
#pragma unroll
    for(int r = 0; r < 3; r++) ti[r] = -(Ai[r * 3] * m[3] + Ai[r * 3 + 1] * m[7] + Ai[r * 3 + 2] * m[11]);
    double KR[9], KRK[9];
    hm::krk_sparse(Ai, (double)fxf, (double)fyf, (double)cxf, (double)cyf, K_inv, KR, KRK);
    const double kt0 = fma((double)cxf, ti[2], (double)fxf * ti[0]), kt1 = fma((double)cyf, ti[2], (double)fyf * ti[1]);
    if(lane == 0)
    {
#pragma unroll
        for(int i = 0; i < 9; i++) payload[12 + i] = (float)KRK[i];
        payload[21] = (float)kt0;
        payload[22] = (float)kt1;
        payload[23] = (float)ti[2];
    }
    __syncwarp();
}

// :348-380 -- digest one so3Step evaluation; returns done
__device__ __noinline__ int solve_so3(Solver * S, const float * s_final, int it)
{
    int done = 0;
    float jtj[9], jtr[3], residual[2];
    hm::unpack_so3(s_final, jtj, jtr, residual);
    S->so3_iterations++;
    float err = sqrtf(residual[0]) / residual[1];                              // :348
    float cnt = residual[1];
    if(err < S->so3_lastError && S->so3_lastCount == cnt) done = 1;            // :352
    else if(err > S->so3_lastError + 0.001)                                     // :356
    {
        err = S->so3_lastError;
        cnt = S->so3_lastCount;
#pragma unroll
        for(int i = 0; i < 9; i++) S->so3_R[i] = S->so3_lastR[i];
        done = 1;
    }
    S->last_so3_error = err;
    S->last_so3_count = cnt;
    if(!done)
    {
        S->so3_lastError = err;
        S->so3_lastCount = cnt;
#pragma unroll
        for(int i = 0; i < 9; i++) S->so3_lastR[i] = S->so3_R[i];
        float delta[3];
        hm::ldlt_solve<float, 3>(jtj, jtr, delta);                               // :368
        // (series form of the exponential map: no sincos / sqrt / division on the single solver thread, ~500 cycles less per
        //  so3Step evaluation; equal to hm::rodrigues to double rounding)
        const double dx = delta[0], dy = delta[1], dz = delta[2];
        double rotUpdate[9];
        rodrigues_small(dx, dy, dz, dx * dx + dy * dy + dz * dz, rotUpdate);
        float ru[9], rl[9];
#pragma unroll
        for(int i = 0; i < 9; i++)
        {
            ru[i] = (float)rotUpdate[i];
            rl[i] = S->so3_R_lr[i];
        }
        hm::mul33(ru, rl, rl);                                                    // :372
#pragma unroll
        for(int i = 0; i < 9; i++)
        {
            S->so3_R_lr[i] = rl[i];
            S->so3_R[i] = rl[i];
        }
        if(it == 10) done = 1; // ten evaluations made
    }
    return done;
}

// :318-329 -- homography K R K^-1 and K R for the next so3Step -> payload in shared memory
__device__ __noinline__ void make_so3_params(const Solver * S, float * payload /*shared*/, int done, float fx, float fy, float cx, float cy,
                                             const double * K_inv)
{
    double R[9], KR[9], H[9];
#pragma unroll
    for(int i = 0; i < 9; i++) R[i] = S->so3_R[i];
    hm::krk_sparse(R, fx, fy, cx, cy, K_inv, KR, H);
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


// Pixels are dealt to the worker CTAs in 32-pixel chunks (one coalesced warp load), round-robin, so that image
// regions with no work (the photometric border, depth holes) are spread evenly and every CTA owns the same number
// of pixels +-32: chunk c -> worker c % W, handled by warp (c / W) % kWarps in pass (c / W) / kWarps.
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

// ------------------------------------------------------------------------------------------------
// photometric candidates of a worker CTA (shared memory, structure of arrays)
//   c0  x | y << 12 | I_next(x,y) << 24        c1  dIdx & 0xffff | dIdy << 16        c2  nextDepth(x,y)
//   r0  u | v << 12 | I_last(u,v) << 24  or kNoMatch                                 r1  lastDepth(u,v)
// ------------------------------------------------------------------------------------------------
struct CandStore
{
    unsigned * c0, * c1;
    float * c2;
    unsigned * r0;
    float * r1;
};

// Everything a pixel's gates read, requested up front and unconditionally (clamped addresses) so that the loads of all
// pixels of a group are in flight together: the level start is otherwise a chain of three L2 round trips per pixel
// (neighbours for the derivative -> depth -> 4x4 window), 24 000 cycles at 640x480.  The four aligned 8-byte runs that hold
// the window rows y-2 .. y+1, columns x-2 .. x+1, also hold the 3x3 stencil of the derivative: nine loads per pixel.
struct CandLoads
{
    int x, y;              // the pixel (0, 0 for an empty slot)
    float d1;              // next depth
    unsigned lo[4], hi[4]; // window rows y-2 .. y+1
    unsigned gxy;          // !DERIV: dIdx | dIdy << 16
};

template<bool DERIV>
__device__ __forceinline__ void candidate_loads(const LevelArgs & L, int u, CandLoads & Q)
{
    const int cols = L.cols, rows = L.rows;
    const int uu = u < 0 ? 0 : u;
    const int y = uu / cols, x = uu - y * cols;
    Q.x = x;
    Q.y = y;
    if constexpr(!DERIV) Q.gxy = ((unsigned)(unsigned short)__ldg(L.dIdx + uu)) | ((unsigned)(unsigned short)__ldg(L.dIdy + uu) << 16);
    Q.d1 = __ldg(L.next_depth + uu);
    const int last = rows * cols - 8;
#pragma unroll
    for(int r = 0; r < 4; r++)
    {
        const int lin = min(max((y + r - 2) * cols + x - 2, 0), last); // no clamp is active where the bytes are used
        const unsigned * wp = reinterpret_cast<const unsigned *>(L.next_image + (lin & ~3));
        Q.lo[r] = __ldg(wp);
        Q.hi[r] = __ldg(wp + 1);
    }
}

// The gates of one pixel from its loads; returns whether it is a candidate and its packed words.  Straight-line code
// (selects, no branches) so that the pixels of a group interleave: with two warps per scheduler a dependent instruction
// costs ~5 cycles and the level start is the length of this instruction stream.  Only pixels on the image border (and
// column 1), where the reference's stencil shifts its taps, take a branch -- to the general statement.
template<bool DERIV>
__device__ __forceinline__ bool candidate_of(const LevelArgs & L, const RgbResParams & RP, int u, const CandLoads & Q, unsigned & w0, unsigned & w1,
                                             float & d1)
{
    const int cols = L.cols, x = Q.x, y = Q.y;
    // bytes[r] = I(y + r - 2, x - 2 .. x + 1), cut out of the aligned run that holds them
    unsigned bytes[4];
#pragma unroll
    for(int r = 0; r < 4; r++) bytes[r] = __funnelshift_r(Q.lo[r], Q.hi[r], 8u * (unsigned)(((y + r - 2) * cols + x - 2) & 3));
    int gx, gy;
    if constexpr(DERIV)
    {
        // computeDerivativeImages (cudafuncs.cu:583-639) fused in: this CTA's pixels of dIdx / dIdy, kept in global
        // memory as well because they are an output of the call (ef_tracker_download) and input of host-solve mode.
        // The stencil of derivative_px (same taps, same order) from the window rows y-1, y, y+1, bytes 1 .. 3:
        const float tl = (float)((bytes[1] >> 8) & 0xffu), tm = (float)((bytes[1] >> 16) & 0xffu), tr = (float)(bytes[1] >> 24);
        const float ml = (float)((bytes[2] >> 8) & 0xffu), mr = (float)(bytes[2] >> 24);
        const float bl = (float)((bytes[3] >> 8) & 0xffu), bm = (float)((bytes[3] >> 16) & 0xffu), br = (float)(bytes[3] >> 24);
        float fx = 0, fy = 0;
        fx = __fmaf_rn(tl, -0.52201f, fx); fx = __fmaf_rn(tr, 0.52201f, fx); fx = __fmaf_rn(ml, -0.79451f, fx);
        fx = __fmaf_rn(mr, 0.79451f, fx);  fx = __fmaf_rn(bl, -0.52201f, fx); fx = __fmaf_rn(br, 0.52201f, fx);
        fy = __fmaf_rn(tl, -0.52201f, fy); fy = __fmaf_rn(tm, -0.79451f, fy); fy = __fmaf_rn(tr, -0.52201f, fy);
        fy = __fmaf_rn(bl, 0.52201f, fy);  fy = __fmaf_rn(bm, 0.79451f, fy);  fy = __fmaf_rn(br, 0.52201f, fy);
        gx = (short)fx;
        gy = (short)fy;
        // valid where no tap left the image and the run of row y+1 was not clamped at the end of the image
        const bool interior = x >= 2 && y >= 1 && x + 1 <= cols - 1 && y + 1 <= L.rows - 1 && (y + 1) * cols + x - 2 <= L.rows * cols - 8;
        if(u >= 0 && !interior)
        {
            short sx, sy;
            const uint8_t * img = L.next_image;
            derivative_px([&](int j, int i) { return __ldg(img + (size_t)j * cols + i); }, L.rows, cols, x, y, sx, sy);
            gx = sx;
            gy = sy;
        }
        if(u >= 0)
        {
            L.dIdx[u] = (short)gx;
            L.dIdy[u] = (short)gy;
        }
    }
    else
    {
        gx = (short)(Q.gxy & 0xffffu);
        gy = (short)(Q.gxy >> 16);
    }
    // the 16-pixel border of RGBResidual (:779-783), the gradient and depth gates, then the 4x4 window
    // [y-2, y+2) x [x-2, x+2) of the next image, which must be non-zero (:787-793)
    const unsigned win = __vcmpne4(bytes[0], 0u) & __vcmpne4(bytes[1], 0u) & __vcmpne4(bytes[2], 0u) & __vcmpne4(bytes[3], 0u);
    const bool keep = u >= 0 && rgb_gate(RP, x, y, gx, gy, Q.d1) && win == 0xffffffffu;
    d1 = Q.d1;
    w0 = (unsigned)x | ((unsigned)y << 12) | (((bytes[2] >> 16) & 0xffu) << 24); // I_next(x, y)
    w1 = ((unsigned)gx & 0xffffu) | ((unsigned)gy << 16);
    return keep;
}

// Level start: the iteration-invariant gates of RGBResidual::getProducts (reduce.cu:779-811) for this CTA's pixels,
// survivors compacted in (pass, warp, lane) order.  kCompactGroup passes are evaluated together (their loads overlap)
// and share one CTA barrier.  Returns the candidate count (uniform over the CTA).
template<bool DERIV>
__device__ __forceinline__ int compact_group(const LevelArgs & L, const RgbResParams & RP, const UnitIter & U, int passes, const CandStore & C,
                                             int * s_wtot /*[2][kCompactGroup][kWarps]*/, int grp, int base)
{
    const unsigned lane = gtid() & 31u, warp = gtid() >> 5;
    const int p0 = grp * kCompactGroup;
    unsigned w0[kCompactGroup], w1[kCompactGroup], ballot[kCompactGroup];
    float d1[kCompactGroup];
    bool keep[kCompactGroup];
    int * wt = s_wtot + (grp & 1) * kCompactGroup * kWarps;
    CandLoads Q[kCompactGroup];
    int un[kCompactGroup];
#pragma unroll
    for(int g = 0; g < kCompactGroup; g++)
    {
        un[g] = (p0 + g < passes) ? U.unit(p0 + g) : -1;
        candidate_loads<DERIV>(L, un[g], Q[g]);
    }
#pragma unroll
    for(int g = 0; g < kCompactGroup; g++) keep[g] = candidate_of<DERIV>(L, RP, un[g], Q[g], w0[g], w1[g], d1[g]);
#pragma unroll
    for(int g = 0; g < kCompactGroup; g++)
    {
        ballot[g] = __ballot_sync(kFullMask, keep[g]);
        if(lane == 0) wt[g * kWarps + warp] = __popc(ballot[g]);
    }
    group_sync();
#pragma unroll
    for(int g = 0; g < kCompactGroup; g++)
    {
        int woff = 0, tot = 0;
#pragma unroll
        for(int w = 0; w < kWarps; w++)
        {
            const int v = wt[g * kWarps + w];
            tot += v;
            woff += (w < (int)warp) ? v : 0;
        }
        if(keep[g])
        {
            const int pos = base + woff + __popc(ballot[g] & ((1u << lane) - 1u));
            C.c0[pos] = w0[g];
            C.c1[pos] = w1[g];
            C.c2[pos] = d1[g];
        }
        base += tot;
    }
    return base;
}

// phase A1: RGBResidual::getProducts past the gates (reduce.cu:813-830) for B of this thread's candidates (rounds
// j0 .. j0 + B of the CTA's list): no control flow, so the B candidates interleave and their gathers are in flight
// together; a slot past the end of the list re-reads candidate 0 and stores nothing
template<int B>
__device__ __forceinline__ void rgb_assoc_batch(const LevelArgs & L, const RgbResParams & RP, const CandStore & C, int n_cand, int j0, int & cnt,
                                                int & sig)
{
    const int cols = L.cols;
    int u0[B], v0[B];
    float td1[B];
    unsigned inext[B];
    bool valid[B], ok[B];
    unsigned q[B]; // (pixel indices fit 32 bits: device_track_supported caps the image at 2^24 pixels)
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        const int c = gtid() + (j0 + k) * kThreads;
        valid[k] = c < n_cand;
        const int cc = valid[k] ? c : 0;
        const unsigned w = C.c0[cc];
        inext[k] = w >> 24;
        ok[k] = rgb_warp(RP, (int)(w & 0xfffu), (int)((w >> 12) & 0xfffu), C.c2[cc], u0[k], v0[k], td1[k]) && valid[k];
        q[k] = ok[k] ? (unsigned)(v0[k] * cols + u0[k]) : 0u;
    }
    float d0s[B];
    unsigned ls[B];
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        d0s[k] = __ldg(L.last_depth + q[k]);
        ls[k] = __ldg(L.last_image + q[k]);
    }
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        const int c = gtid() + (j0 + k) * kThreads;
        const bool good = ok[k] && rgb_accept(RP, td1[k], d0s[k], (uint8_t)ls[k]);
        const float diff = static_cast<float>(inext[k]) - static_cast<float>(ls[k]); // reduce.cu:827
        const unsigned rec = good ? ((unsigned)u0[k] | ((unsigned)v0[k] << 12) | (ls[k] << 24)) : kNoMatch;
        if(valid[k])
        {
            C.r0[c] = rec;
            C.r1[c] = d0s[k];
        }
        cnt += good ? 1 : 0;
        sig += good ? (int)(diff * diff) : 0; // reduce.cu:830
    }
}

__device__ __forceinline__ void rgb_assoc_cands(const LevelArgs & L, const RgbResParams & RP, const CandStore & C, int n_cand, int & cnt, int & sig)
{
    const int rounds = (n_cand + kThreads - 1) / kThreads; // uniform over the CTA
    int j = 0;
    for(; j + kRgbBatch <= rounds; j += kRgbBatch) rgb_assoc_batch<kRgbBatch>(L, RP, C, n_cand, j, cnt, sig);
    if(kRgbBatch > 2 && rounds - j == 2) rgb_assoc_batch<2>(L, RP, C, n_cand, j, cnt, sig);
    else if(kRgbBatch > 1 && rounds - j == 1) rgb_assoc_batch<1>(L, RP, C, n_cand, j, cnt, sig);
}

// phase B: RGBReduction::getProducts (reduce.cu:512-595) for B of this thread's candidates, again without control
// flow: unmatched candidates contribute rows of exact zeros
template<int B>
__device__ __forceinline__ bool rgb_rows_batch(const RgbStepParams & SP, const CandStore & C, int n_cand, int j0, float (*accRs)[32])
{
    bool any = false;
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        float * accR = accRs[k % kSets]; // j0 is a multiple of kSets: round j0 + k belongs to virtual thread set k % kSets
        const int c = gtid() + (j0 + k) * kThreads;
        const int cc = c < n_cand ? c : 0;
        const unsigned rec = C.r0[cc];
        const bool good = c < n_cand && rec != kNoMatch;
        const unsigned w = C.c0[cc], g = C.c1[cc];
        const float Z = C.r1[cc];
        const float diff = static_cast<float>(w >> 24) - static_cast<float>(rec >> 24);
        const float3 cp = project_point((int)(rec & 0xfffu), (int)((rec >> 12) & 0xfffu), Z, SP.inv_fx, SP.inv_fy, SP.cx, SP.cy);
        float row[7];
        rgb_row(SP, diff, cp.x, cp.y, cp.z, (short)(g & 0xffffu), (short)(g >> 16), row);
#pragma unroll
        for(int i = 0; i < 7; i++) row[i] = good ? row[i] : 0.f;
        int kk = 0;
#pragma unroll
        for(int i = 0; i < 6; i++)
        {
#pragma unroll
            for(int j = i; j < 7; j++) accR[kk++] += row[i] * row[j];
        }
        accR[27] += row[6] * row[6];
        accR[28] += good ? 1.0f : 0.f;
        any |= good;
    }
    return any;
}

__device__ __forceinline__ bool rgb_rows_cands(const RgbStepParams & SP, const CandStore & C, int n_cand, float (*accR)[32])
{
    const int rounds = (n_cand + kThreads - 1) / kThreads;
    bool any = false;
    int j = 0;
    for(; j + kRgbBatch <= rounds; j += kRgbBatch) any |= rgb_rows_batch<kRgbBatch>(SP, C, n_cand, j, accR);
    if(kRgbBatch > 2 && rounds - j == 2) any |= rgb_rows_batch<2>(SP, C, n_cand, j, accR);
    else if(kRgbBatch > 1 && rounds - j == 1) any |= rgb_rows_batch<1>(SP, C, n_cand, j, accR);
    return any;
}

// Level start: the current-frame vertex / normal of this CTA's pixels -> shared memory (they do not change over the
// iterations of a level; 24 B per pixel, structure of arrays, slot = pass * kThreads + thread)
constexpr int kStageBatch = EF_TRACK_STAGE_BATCH; // passes of a thread whose six loads each are requested together when the current maps are staged
__device__ __forceinline__ void stage_batch(const LevelArgs & L, const UnitIter & U, int passes, float * s_vn, int cap, int p0)
{
    // one batch = one L2 round trip; nothing else is live in the registers at this point
    const unsigned plane = (unsigned)(L.rows * L.cols);
    float v[kStageBatch][6];
    int u[kStageBatch];
#pragma unroll
    for(int k = 0; k < kStageBatch; k++)
    {
        u[k] = (p0 + k < passes) ? U.unit(p0 + k) : -1;
        const int uu = u[k] >= 0 ? u[k] : 0;
        v[k][0] = __ldg(L.vc + uu); v[k][1] = __ldg(L.vc + plane + uu); v[k][2] = __ldg(L.vc + 2 * plane + uu);
        v[k][3] = __ldg(L.nc + uu); v[k][4] = __ldg(L.nc + plane + uu); v[k][5] = __ldg(L.nc + 2 * plane + uu);
    }
#pragma unroll
    for(int k = 0; k < kStageBatch; k++)
        if(u[k] >= 0)
        {
            float * q = s_vn + (p0 + k) * kThreads + gtid();
#pragma unroll
            for(int c = 0; c < 6; c++) q[c * cap] = v[k][c];
        }
}

// phase A2: ICPReduction (reduce.cu:285-347) for up to B of this thread's pixels (passes p0 .. p0 + B): all the
// coalesced loads first, then the projections, then all the gathers, then the products
template<int B, bool SMEM>
__device__ __forceinline__ bool icp_batch(const LevelArgs & L, const IcpParams & IP, const UnitIter & U, int p0, int passes, float (*accIs)[32],
                                          const float * s_vn, int cap)
{
    const int cols = L.cols;
    const unsigned plane = (unsigned)(L.rows * cols); // 32-bit indices: one IMAD.WIDE per gather instead of 64-bit adds
    float3 v[B], n[B];
    bool in1[B];
    bool any = false;
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        const int uu = (p0 + k < passes) ? U.unit(p0 + k) : -1;
        in1[k] = uu >= 0;
        any |= in1[k];
        const int u = in1[k] ? uu : 0;
        if constexpr(SMEM)
        {
            // iteration-invariant: staged once per level (stage_current_maps), slot = pass * kThreads + thread
            const float * sq = s_vn + (in1[k] ? (p0 + k) * kThreads + gtid() : 0);
            v[k].x = sq[0]; v[k].y = sq[cap]; v[k].z = sq[2 * cap];
            n[k].x = sq[3 * cap]; n[k].y = sq[4 * cap]; n[k].z = sq[5 * cap];
        }
        else
        {
            v[k].x = L.vc[u]; v[k].y = L.vc[plane + u]; v[k].z = L.vc[2 * plane + u];
            n[k].x = L.nc[u]; n[k].y = L.nc[plane + u]; n[k].z = L.nc[2 * plane + u];
        }
    }
    // from here on no control flow: the B pixels interleave (two warps per scheduler need the instruction-level
    // parallelism); rejected pixels gather pixel 0 and contribute rows of exact zeros
    float3 vg[B], vp[B], np[B];
    unsigned q[B];
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        int ux, uy;
        in1[k] = icp_project(IP, v[k], vg[k], ux, uy) && in1[k];
        q[k] = in1[k] ? (unsigned)(uy * cols + ux) : 0u;
    }
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        vp[k].x = __ldg(L.vp + q[k]); vp[k].y = __ldg(L.vp + plane + q[k]); vp[k].z = __ldg(L.vp + 2 * plane + q[k]);
        np[k].x = __ldg(L.np + q[k]); np[k].y = __ldg(L.np + plane + q[k]); np[k].z = __ldg(L.np + 2 * plane + q[k]);
    }
#pragma unroll
    for(int k = 0; k < B; k++)
    {
        float row[7];
        float * accI = accIs[k % kSets]; // p0 is a multiple of kSets: pass p0 + k belongs to virtual warp set k % kSets
        const bool ok = icp_finish_select(IP, vg[k], n[k], vp[k], np[k], row) && in1[k];
#pragma unroll
        for(int i = 0; i < 7; i++) row[i] = ok ? row[i] : 0.f;
        int kk = 0;
#pragma unroll
        for(int i = 0; i < 6; i++)
        {
#pragma unroll
            for(int j = i; j < 7; j++) accI[kk++] += row[i] * row[j];
        }
        accI[27] += row[6] * row[6];
        accI[28] += ok ? 1.0f : 0.f;
    }
    return any;
}

// all ICP passes [p_begin, p_end) of this thread in batches of kIcpBatch; the remainder runs in a batch of its own
// size so that no empty pixel slot is computed (p_begin, p_end are uniform over the CTA).  (Software-pipelining the
// batches -- gathers of batch k+1 issued before batch k is finished -- was measured: no gain in the ICP phase, which
// is bound by instruction issue with two warps per scheduler, and the 255 registers it needs slow the other phases.)
template<bool SMEM>
__device__ __forceinline__ bool icp_passes(const LevelArgs & L, const IcpParams & IP, const UnitIter & U, int p_begin, int p_end, float (*accI)[32],
                                           const float * s_vn, int cap)
{
    static_assert(kIcpBatch >= 1 && kIcpBatch <= 4 && kIcpBatch % kSets == 0, "EF_TRACK_ICP_BATCH (a multiple of EF_TRACK_SETS)");
    static_assert(kRgbBatch % kSets == 0, "EF_TRACK_RGB_BATCH (a multiple of EF_TRACK_SETS)");
    bool any = false;
    int p = p_begin;
    for(; p + kIcpBatch <= p_end; p += kIcpBatch) any |= icp_batch<kIcpBatch, SMEM>(L, IP, U, p, p_end, accI, s_vn, cap);
    const int rem = p_end - p;
    if(rem == 3) any |= icp_batch<3, SMEM>(L, IP, U, p, p_end, accI, s_vn, cap);
    else if(rem == 2) any |= icp_batch<2, SMEM>(L, IP, U, p, p_end, accI, s_vn, cap);
    else if(rem == 1) any |= icp_batch<1, SMEM>(L, IP, U, p, p_end, accI, s_vn, cap);
    return any;
}

// RGBDOdometry.cpp:461-462 from the barrier-B sums (the precedence quirk of :461 is kept)
__device__ __forceinline__ void sigma_from_sums(int sigma, int rgbSize, float & sigmaVal, float & rgbError)
{
    sigmaVal = (float)sqrt((double)((((float)sigma / (float)rgbSize) == 0) ? 1 : rgbSize));
    rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
}

// static shared memory of one thread group
struct GroupShared
{
    float red[kVWarps * 64];
    float final_[64];
    float par[2][kPayload];
    float sigma[3];                // sigmaVal, rgbError, (float)rgbSize of the current iteration
    int wcnt[kWarps], wsig[kWarps];
    int wtot[2 * kCompactGroup * kWarps];
    int flag;
    alignas(16) Solver solver;
};

struct BatchArgs
{
    TrackArgs seq[kGroups];
};

#if !defined(EF_TRACK_ALT) && EF_TRACK_GROUPS == 1
// ================================================================================================
// The symmetric body (single launches on a SUBSET of the SMs, EF_OPT_GRID_CTAS): EVERY CTA is a worker AND solves.
//
// Round 1's structure -- CTA 0 gathers the rows, solves, publishes the parameters, 147 workers wait -- costs two L2 hops per
// iteration (rows -> CTA 0, parameters -> workers: 4 400 cycles of communication before the 1 900 of the solve,
// tools/gather2_probe.cu).  Here every CTA polls ALL rows itself, adds them in row order and runs the solve redundantly:
// one hop (2 980 cycles in the same probe), identical bits in every CTA because the instructions and their inputs are identical,
// no parameter line, no solver CTA.  Barrier B works the same way: every CTA reads all {count, sum} arrivals and forms the
// robust-weight scale itself.  CTA 0 alone stores the result.
// Measured (profiles/r02_k_track_experiments.txt): on the whole GPU this body does NOT beat track_body -- the redundant FP64 solve is
// slower when every SM issues it (2 400 vs 1 890 cycles) and the level preparation, which track_body's workers do while they wait for
// parameters, has no wait to hide in (171.5 -> 174.6 us) -- but a handle that owns a quarter of the SMs never had room for the
// three levels' lists side by side, so its preparation was in the open already and the saved hop is a net gain (4 x 37 SMs:
// 8 200 -> 8 700 frames/s).  ef_track_dispatch / device_track_configure therefore pick it for partitioned handles only.
//   rows    G x kRowChunks flagged chunks, two copies by the parity of the round (a CTA may publish round n + 1 while a slower
//           one still reads round n; n + 2 needs that CTA's row of n + 1, which it sends after it is done reading n)
//   bslot   G flagged chunks, two copies by the parity of the iteration, same argument
// ================================================================================================
template<bool TIMING>
__device__ __forceinline__ void track_body_sym(const TrackArgs & A, GroupShared & GS, int4 * const s_dyn, const int b, const int G)
{
    float * const s_red = GS.red;
    float * const s_final = GS.final_;
    float (* const s_par)[kPayload] = GS.par;
    float * const s_sigma = GS.sigma;
    int * const s_wcnt = GS.wcnt, * const s_wsig = GS.wsig;
    int * const s_wtot = GS.wtot;
    int & s_flag = GS.flag;
    Solver * const S = &GS.solver;

    const int W = G, widx = b;                  // every CTA owns a share of the pixels
    const bool is_writer = (b == 0) && gtid() == 0;
    const unsigned lane = gtid() & 31u, warp = gtid() >> 5;
    float * const s_rows = reinterpret_cast<float *>(s_dyn); // the gathered rows, G x kRowFloats floats; the lists follow at A.cand_base
    char * const s_lists = reinterpret_cast<char *>(s_dyn) + A.cand_base;
    auto cand_store = [&](int lv) {
        CandStore c;
        unsigned * base = reinterpret_cast<unsigned *>(s_lists + A.lvl_off[lv]);
        const int cap = A.lvl_cap[lv];
        c.c0 = base;
        c.c1 = base + cap;
        c.c2 = reinterpret_cast<float *>(base + 2 * cap);
        c.r0 = base + 3 * cap;
        c.r1 = reinterpret_cast<float *>(base + 4 * cap);
        return c;
    };
    auto vn_store = [&](int lv) { return reinterpret_cast<float *>(s_lists + A.lvl_off[lv]) + 5 * (size_t)A.lvl_cap[lv]; };
    const size_t row_copy = (size_t)kMaxGrid * kRowChunks, slot_copy = kSlotCopy; // chunks between the two copies
    auto rows_of = [&](unsigned round) { return A.rows + (round & 1u) * row_copy; };
    auto slots_of = [&](unsigned it) { return A.bslot + (it & 1u) * slot_copy; };

    unsigned rel = A.epoch_base; // iterations started (barrier-B arrivals)
    unsigned arr = A.epoch_base; // row rounds

    if(gtid() == 0)
    {
#pragma unroll
        for(int i = 0; i < 16; i++) S->resultRt[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
        for(int i = 0; i < 9; i++) S->Rcurr[i] = A.Rprev[i];
#pragma unroll
        for(int i = 0; i < 3; i++) S->tcurr[i] = A.tprev[i];
#pragma unroll
        for(int i = 0; i < 9; i++) S->Rprev[i] = A.Rprev[i];
#pragma unroll
        for(int i = 0; i < 3; i++) S->tprev[i] = A.tprev[i];
        S->last_icp_error = A.prev_icp_error; S->last_icp_count = A.prev_icp_count;
        S->last_so3_error = A.prev_so3_error; S->last_so3_count = A.prev_so3_count;
        S->last_rgb_error = A.prev_rgb_error; S->last_rgb_count = A.prev_rgb_count;
        S->so3_iterations = 0;
        S->se3_iterations[0] = S->se3_iterations[1] = S->se3_iterations[2] = 0;
    }
    group_sync();

    int dbg_it = 0;
    auto stamp = [&](int k) {
        if constexpr(TIMING)
        {
            if(gtid() == 0 && dbg_it < kMaxIters) A.dbg[((size_t)blockIdx.x * kMaxIters + dbg_it) * kDbgStamps + k] = clock64();
        }
    };
    if constexpr(TIMING)
    {
        if(is_writer) A.dbg[(size_t)(kMaxIters - 2) * kDbgStamps] = clock64();
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
        if(gtid() == 0)
        {
#pragma unroll
            for(int i = 0; i < 9; i++)
            {
                S->so3_R[i] = S->so3_lastR[i] = (i % 4 == 0) ? 1.0 : 0.0;
                S->so3_R_lr[i] = (i % 4 == 0) ? 1.f : 0.f;
            }
            S->so3_lastError = FLT_MAX / 2;
            S->so3_lastCount = FLT_MAX / 2;
        }
        UnitIter U{L.rows * L.cols, W, widx, (int)lane, (int)warp};
        const int passes = U.passes();

        for(int it = 0; it <= 10; it++)
        {
            ++rel;
            // ---- everybody: digest the previous evaluation (:348-380), derive the next homography ----
            if(it > 0) gather_rows(rows_of(arr), G, kSo3Chunks, arr, s_rows, s_red, s_final);
            float * par = s_par[rel & 1u];
            if(gtid() == 0)
            {
                int d = 0;
                if(it > 0) d = solve_so3(S, s_final, it);
                make_so3_params(S, par, d, L.fx, L.fy, L.cx, L.cy, L.K_inv);
                s_flag = d;
            }
            group_sync();
            if(s_flag != 0) break;
            P.image_basis = mat_from(par);
            P.krlr = mat_from(par + 9);
            ++arr;

            // ---- so3Step over this CTA's pixels ----
            float acc[kSets][16];
#pragma unroll
            for(int q = 0; q < kSets; q++)
#pragma unroll
                for(int i = 0; i < 16; i++) acc[q][i] = 0.f;
            for(int p = 0; p < passes; p += kSets)
            {
#pragma unroll
                for(int q = 0; q < kSets; q++) // pass p + q belongs to virtual warp set q (p is a multiple of kSets)
                {
                    const int u = (p + q < passes) ? U.unit(p + q) : -1;
                    if(u >= 0)
                    {
                        const int y = u / L.cols, x = u - y * L.cols;
                        float row[4];
                        if(so3_row(P, x, y, A.so3_last, A.so3_next, L.cols, row)) accumulate_so3(acc[q], row);
                    }
                }
            }
#pragma unroll
            for(int q = 0; q < kSets; q++)
            {
                const float lane_value = warp_transpose_reduce16(acc[q]);
                if(lane < 16) s_red[(warp + q * kWarps) * 16 + lane] = lane_value;
            }
            group_sync();
            if(gtid() < 16)
            {
                float s = 0.f;
#pragma unroll
                for(int w = 0; w < kVWarps; w++) s += s_red[w * 16 + gtid()];
                s_final[gtid()] = (gtid() < 11) ? s : 0.f;
            }
            group_sync();
            publish_row(rows_of(arr) + (size_t)widx * kSo3Chunks, s_final, kSo3Chunks, arr);
        }
        if(gtid() == 0)
        {
#pragma unroll
            for(int x = 0; x < 3; x++)
            {
#pragma unroll
                for(int y = 0; y < 3; y++) S->resultRt[x * 4 + y] = S->so3_R[x * 3 + y]; // :394-403
            }
        }
        group_sync();
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
    bool pending = false; // a row round nobody has digested yet
    int pending_level = 0;

    // Level preparation in units (one compaction group or one staging batch each, ~5 000 cycles): a level's units run at its
    // start at the latest, but when all levels' lists fit shared memory side by side (A.prework) they run earlier, one per
    // iteration of a coarser level, in the window where the rows of the other CTAs are still on their way.
    int prep_unit[kNumPyrs] = {0, 0, 0}, prep_cand[kNumPyrs] = {0, 0, 0};
    auto prep_passes = [&](int lv) { return UnitIter{A.lvl[lv].rows * A.lvl[lv].cols, W, widx, (int)lane, (int)warp}.passes(); };
    auto prep_groups = [&](int lv) { return A.rgb ? (prep_passes(lv) + kCompactGroup - 1) / kCompactGroup : 0; };
    auto prep_units = [&](int lv) {
        return prep_groups(lv) + ((A.icp && A.icp_in_smem) ? (prep_passes(lv) + kStageBatch - 1) / kStageBatch : 0);
    };
    auto prepare_unit = [&](int lv) {
        const LevelArgs & PL = A.lvl[lv];
        const UnitIter PU{PL.rows * PL.cols, W, widx, (int)lane, (int)warp};
        const int pp = PU.passes(), groups = prep_groups(lv), unit = prep_unit[lv]++;
        if(unit < groups)
        {
            RgbResParams PR;
            PR.min_scale = PL.min_scale;
            PR.max_depth_delta = A.max_depth_delta;
            PR.rows = PL.rows;
            PR.cols = PL.cols;
            const CandStore PC = cand_store(lv);
            prep_cand[lv] = A.make_derivatives ? compact_group<true>(PL, PR, PU, pp, PC, s_wtot, unit, prep_cand[lv])
                                               : compact_group<false>(PL, PR, PU, pp, PC, s_wtot, unit, prep_cand[lv]);
        }
        else
            stage_batch(PL, PU, pp, vn_store(lv), A.lvl_cap[lv], (unit - groups) * kStageBatch);
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
        UnitIter U{L.rows * L.cols, W, widx, (int)lane, (int)warp};
        const int passes = U.passes();

        // ---- level start: whatever is left of this level's preparation (candidate list, staged current maps) ----
        const CandStore C = cand_store(lv);
        float * const s_vn = vn_store(lv);
        dbg_it = it_global;
        stamp(10);
        if(!A.prework) group_sync(); // the levels share one region: every thread is done with the previous level's records
        while(prep_unit[lv] < prep_units(lv)) prepare_unit(lv);
        group_sync();
        stamp(11);
        stamp(12);
        const int n_cand = prep_cand[lv];
        int next_lv = lv - 1;
        while(next_lv >= 0 && A.lvl[next_lv].iterations <= 0) next_lv--;

        float lastRGBError = FLT_MAX;      // every thread tracks it for the uniform rgb_only break (:464)
        bool first_of_level = true;

        for(int j = 0; j < L.iterations; j++)
        {
            const int cnt_slot = it_global++;
            dbg_it = cnt_slot;
            stamp(0);
            ++rel;
            // a unit of the next level's preparation while the other CTAs' rows are on their way
            if(A.prework && j > 0)
            {
                while(next_lv >= 0 && prep_unit[next_lv] >= prep_units(next_lv))
                {
                    next_lv--;
                    while(next_lv >= 0 && A.lvl[next_lv].iterations <= 0) next_lv--;
                }
                if(next_lv >= 0) prepare_unit(next_lv);
            }
            // ---- everybody: finish the previous iteration (:541-583) and derive this one's pose (:480-481) and photometric
            //      warp (:424-434) -- warps 0 and 1 side by side, identical in every CTA ----
            const bool was_pending = pending;
            if(pending)
            {
                gather_rows(rows_of(arr), G, kRowChunks, arr, s_rows, s_red, s_final);
                stamp(2);
                if(warp == 0)
                    warp_solve_se3(S, s_final, A.icp, A.rgb, A.icp_weight, pending_level, A.rgb != 0,
                                   (TIMING && b == 0 && dbg_it < kMaxIters) ? A.dbg + (size_t)dbg_it * kDbgStamps : nullptr);
                stamp(3);
            }
            if(gtid() == 0 && first_of_level) S->last_rgb_error = FLT_MAX; // :420
            float * par = s_par[rel & 1u];
            if(warp == 0)
            {
                // (with a solve pending the hand-over to warp 1 happened inside warp_solve_se3, right after the SE(3) update)
                if(!was_pending && A.rgb) solver_pair_sync();
                warp_make_pose(S, par);
                stamp(7);
            }
            else if(warp == 1 && A.rgb)
            {
                solver_pair_sync();
                warp_make_rgb_params(S, par, L.fx, L.fy, L.cx, L.cy, L.K_inv);
            }
            group_sync();
            stamp(1);
            stamp(4);
            IP.Rcurr = mat_from(par);
            IP.tcurr = make_float3(par[9], par[10], par[11]);
            RP.krkinv = mat_from(par + 12);
            RP.kt = make_float3(par[21], par[22], par[23]);
            pending = false;
            first_of_level = false;

            // ---- phase A1: photometric association of the candidates -> records in shared memory, then the barrier-B arrival:
            //      this CTA's {count, sum diff^2} as one flagged chunk (integer sums are exact in any order and sigma wraps
            //      mod 2^32 exactly like the reference's int sum) ----
            if(A.rgb)
            {
                int cnt = 0, sig = 0;
                rgb_assoc_cands(L, RP, C, n_cand, cnt, sig);
                cnt = __reduce_add_sync(kFullMask, cnt);
                sig = __reduce_add_sync(kFullMask, sig);
                if(lane == 0) { s_wcnt[warp] = cnt; s_wsig[warp] = sig; }
                group_sync();
                if(gtid() == 0)
                {
                    unsigned c = 0, s = 0;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) { c += (unsigned)s_wcnt[w]; s += (unsigned)s_wsig[w]; }
                    st_relaxed_v4(slots_of(rel) + widx, make_uint4(c, s, 0u, rel));
                }
            }
            stamp(5);

            // ---- phase A2: ICP association + 29 sums (hides the barrier-B hop and the inter-CTA skew) ----
            float accI[kSets][32];
#pragma unroll
            for(int q = 0; q < kSets; q++)
#pragma unroll
                for(int i = 0; i < 32; i++) accI[q][i] = 0.f;
            {
                float vi[kSets];
#pragma unroll
                for(int q = 0; q < kSets; q++) vi[q] = 0.f;
                if(A.icp)
                {
                    const bool anyI = A.icp_in_smem ? icp_passes<true>(L, IP, U, 0, passes, accI, s_vn, A.lvl_cap[lv])
                                                    : icp_passes<false>(L, IP, U, 0, passes, accI, s_vn, A.lvl_cap[lv]);
                    if(__any_sync(kFullMask, anyI)) // a warp without pixels contributes zeros
                    {
#pragma unroll
                        for(int q = 0; q < kSets; q++) vi[q] = warp_transpose_reduce32(accI[q]);
                    }
                }
#pragma unroll
                for(int q = 0; q < kSets; q++)
                    if(lane < 29) s_red[(warp + q * kWarps) * 64 + lane] = vi[q];
                stamp(8);
            }

            // ---- barrier B: every CTA reads all arrivals, adds them and forms the robust-weight scale (:461-462) itself ----
            bool level_break = false;
            if(A.rgb)
            {
                int c = 0, sg = 0;
                const uint4 * slots = slots_of(rel);
                for(int w = gtid(); w < G; w += kThreads)
                {
                    uint4 v;
                    do
                    {
                        v = ld_relaxed_v4(slots + w);
                    } while(v.w != rel);
                    c += (int)v.x;
                    sg += (int)v.y;
                }
                c = __reduce_add_sync(kFullMask, c);
                sg = __reduce_add_sync(kFullMask, sg);
                group_sync(); // (s_wcnt / s_wsig: everybody is past the CTA total of phase A1)
                if(lane == 0) { s_wcnt[warp] = c; s_wsig[warp] = sg; }
                group_sync();
                if(gtid() == 0)
                {
                    unsigned cc = 0, ss = 0;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) { cc += (unsigned)s_wcnt[w]; ss += (unsigned)s_wsig[w]; }
                    float sigmaVal, rgbError;
                    sigma_from_sums((int)ss, (int)cc, sigmaVal, rgbError); // :461-462
                    s_sigma[0] = sigmaVal;
                    s_sigma[1] = rgbError;
                    s_sigma[2] = (float)(int)cc;
                }
                group_sync();
                stamp(6);
                const float rgbError = s_sigma[1];
                if(A.rgb_only && rgbError > lastRGBError) level_break = true; // :464 (uniform across the grid)
                if(!level_break)
                {
                    lastRGBError = rgbError;
                    if(gtid() == 0)
                    {
                        S->last_rgb_error = rgbError; // :469-470
                        S->last_rgb_count = s_sigma[2];
                    }
                    SP.sigma = A.rgb_only ? -1.f : s_sigma[0]; // :472-475
                }
            }
            else if(gtid() == 0)
            {
                // :461-470 run even without RGB: sigma = rgbSize = 0 -> rgbError 0, count 0
                S->last_rgb_error = 0.f;
                S->last_rgb_count = 0.f;
            }
            if(level_break) break; // no row outstanding: every CTA takes the same branch

            ++arr;
            pending = true;
            pending_level = lv;

            // ---- phase B: photometric rows from the records in shared memory -> 29 more sums; this CTA's row of the round ----
            {
                float vr[kSets];
#pragma unroll
                for(int q = 0; q < kSets; q++) vr[q] = 0.f;
                if(A.rgb)
                {
                    float accR[kSets][32];
#pragma unroll
                    for(int q = 0; q < kSets; q++)
#pragma unroll
                        for(int i = 0; i < 32; i++) accR[q][i] = 0.f;
                    const bool any = rgb_rows_cands(SP, C, n_cand, accR);
                    if(__any_sync(kFullMask, any))
                    {
#pragma unroll
                        for(int q = 0; q < kSets; q++) vr[q] = warp_transpose_reduce32(accR[q]);
                    }
                }
#pragma unroll
                for(int q = 0; q < kSets; q++)
                    if(lane < 29) s_red[(warp + q * kWarps) * 64 + 29 + lane] = vr[q];
                group_sync();
                if(gtid() < kRowFloats)
                {
                    float sum = 0.f;
                    if(gtid() < 58)
                    {
#pragma unroll
                        for(int w = 0; w < kVWarps; w++) sum += s_red[w * 64 + gtid()];
                    }
                    s_final[gtid()] = sum;
                }
                group_sync();
                publish_row(rows_of(arr) + (size_t)widx * kRowChunks, s_final, kRowChunks, arr);
                stamp(9);
            }
        }
    }

    // ============================================================================================
    // epilogue: last solve (every CTA; CTA 0 writes), jump rejection (:587-591), outputs
    // ============================================================================================
    if(b != 0) return; // (their last rows stay where they are: nothing rewrites them in this launch)
    if(pending)
    {
        gather_rows(rows_of(arr), G, kRowChunks, arr, s_rows, s_red, s_final);
        if(warp == 0) warp_solve_se3(S, s_final, A.icp, A.rgb, A.icp_weight, pending_level, false, nullptr);
        group_sync();
    }
    if(is_writer)
    {
        float Rc[9], tc[3];
#pragma unroll
        for(int i = 0; i < 9; i++) Rc[i] = S->Rcurr[i];
#pragma unroll
        for(int i = 0; i < 3; i++) tc[i] = S->tcurr[i];
        if(A.rgb)
        {
            const float d[3] = {tc[0] - A.tprev[0], tc[1] - A.tprev[1], tc[2] - A.tprev[2]};
            if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
            {
#pragma unroll
                for(int i = 0; i < 9; i++) Rc[i] = A.Rprev[i];
#pragma unroll
                for(int i = 0; i < 3; i++) tc[i] = A.tprev[i];
            }
        }
        TrackOutput * out = A.out;
#pragma unroll
        for(int i = 0; i < 3; i++) out->trans[i] = tc[i];
#pragma unroll
        for(int i = 0; i < 9; i++) out->rot[i] = Rc[i];
        out->st.last_icp_error = S->last_icp_error; out->st.last_icp_count = S->last_icp_count;
        out->st.last_rgb_error = S->last_rgb_error; out->st.last_rgb_count = S->last_rgb_count;
        out->st.last_so3_error = S->last_so3_error; out->st.last_so3_count = S->last_so3_count;
        out->st.so3_iterations = S->so3_iterations;
#pragma unroll
        for(int i = 0; i < 3; i++) out->st.se3_iterations[i] = S->se3_iterations[i];
        if(S->se3_iterations[0] + S->se3_iterations[1] + S->se3_iterations[2] > 0)
        {
            // lastA / lastb of the last solve (reduce.cu:475-486 unpack order)
#pragma unroll
            for(int i = 0; i < 6; i++)
            {
#pragma unroll
                for(int j = i; j < 7; j++)
                {
                    const double v = S->last_S[hm::acc_index(i, j)];
                    if(j == 6) out->st.last_b[i] = v;
                    else out->st.last_A[j * 6 + i] = out->st.last_A[i * 6 + j] = v;
                }
            }
        }
        // the host polls `status` in pinned memory (device_track_finish): result first, fence, then the flag.
        // 2 = no solve ran: lastA / lastb keep their previous values (host side)
        const int status = (S->se3_iterations[0] + S->se3_iterations[1] + S->se3_iterations[2] > 0) ? 1 : 2;
        if constexpr(TIMING) A.dbg[(size_t)(kMaxIters - 1) * kDbgStamps] = clock64();
        __threadfence_system();
        *reinterpret_cast<volatile int *>(&out->status) = status;
        __threadfence_system();
    }
}

template<bool TIMING>
__global__ void __launch_bounds__(kThreads, 1) k_track_sym(const __grid_constant__ BatchArgs BA)
{
    extern __shared__ int4 s_dyn_sym[];
    __shared__ GroupShared s_group;
    track_body_sym<TIMING>(BA.seq[0], s_group, s_dyn_sym, (int)blockIdx.x, (int)gridDim.x);
}
#endif // symmetric body

// One sequence of a launch, as seen by one thread group: the solver CTA of the sequence (is_solver_cta: gathers, solves, publishes)
// or worker `widx` of its W workers.  s_dyn: the group's dynamic shared memory -- workers: candidates + match records; solver:
// the gathered rows.
template<bool TIMING>
__device__ __forceinline__ void track_body(const TrackArgs & A, GroupShared & GS, int4 * const s_dyn, const bool is_solver_cta, const int widx, const int W)
{
    float * const s_red = GS.red;
    float * const s_final = GS.final_;
    float (* const s_par)[kPayload] = GS.par;
    float * const s_sigma = GS.sigma;
    int * const s_wcnt = GS.wcnt, * const s_wsig = GS.wsig;
    int * const s_wtot = GS.wtot;
    int & s_flag = GS.flag;
    Solver & s_solver = GS.solver;

    const uint4 * my_par = A.par + (size_t)(widx % kReplicas) * kReplicaStride;
    const bool is_solver = is_solver_cta && gtid() == 0;
    const unsigned lane = gtid() & 31u, warp = gtid() >> 5;
    float * s_rows = reinterpret_cast<float *>(s_dyn);
    // per level: five candidate arrays of lvl_cap entries, then (icp_in_smem) six planes of lvl_cap floats
    auto cand_store = [&](int lv) {
        CandStore c;
        unsigned * base = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(s_dyn) + A.lvl_off[lv]);
        const int cap = A.lvl_cap[lv];
        c.c0 = base;
        c.c1 = base + cap;
        c.c2 = reinterpret_cast<float *>(base + 2 * cap);
        c.r0 = base + 3 * cap;
        c.r1 = reinterpret_cast<float *>(base + 4 * cap);
        return c;
    };
    auto vn_store = [&](int lv) { return reinterpret_cast<float *>(reinterpret_cast<char *>(s_dyn) + A.lvl_off[lv]) + 5 * (size_t)A.lvl_cap[lv]; };
    Solver * S = &s_solver;

    // epochs, tracked identically by every thread of the grid
    unsigned rel = A.epoch_base; // parameter publications
    unsigned arr = A.epoch_base; // row rounds

    if(is_solver)
    {
#pragma unroll
        for(int i = 0; i < 16; i++) S->resultRt[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
        for(int i = 0; i < 9; i++) S->Rcurr[i] = A.Rprev[i];
#pragma unroll
        for(int i = 0; i < 3; i++) S->tcurr[i] = A.tprev[i];
#pragma unroll
        for(int i = 0; i < 9; i++) S->Rprev[i] = A.Rprev[i];
#pragma unroll
        for(int i = 0; i < 3; i++) S->tprev[i] = A.tprev[i];
        S->last_icp_error = A.prev_icp_error; S->last_icp_count = A.prev_icp_count;
        S->last_so3_error = A.prev_so3_error; S->last_so3_count = A.prev_so3_count;
        S->last_rgb_error = A.prev_rgb_error; S->last_rgb_count = A.prev_rgb_count;
        S->so3_iterations = 0;
        S->se3_iterations[0] = S->se3_iterations[1] = S->se3_iterations[2] = 0;
    }

    int dbg_it = 0;
    auto stamp = [&](int k) {
        if constexpr(TIMING)
        {
            if(gtid() == 0 && dbg_it < kMaxIters) A.dbg[((size_t)blockIdx.x * kMaxIters + dbg_it) * kDbgStamps + k] = clock64();
        }
    };
    // kernel start / end of CTA 0 live in the two last iteration slots (a call runs at most 19 SE3 iterations)
    if constexpr(TIMING)
    {
        if(is_solver) A.dbg[(size_t)(kMaxIters - 2) * kDbgStamps] = clock64();
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
        if(is_solver)
        {
#pragma unroll
            for(int i = 0; i < 9; i++)
            {
                S->so3_R[i] = S->so3_lastR[i] = (i % 4 == 0) ? 1.0 : 0.0;
                S->so3_R_lr[i] = (i % 4 == 0) ? 1.f : 0.f;
            }
            S->so3_lastError = FLT_MAX / 2;
            S->so3_lastCount = FLT_MAX / 2;
        }
        UnitIter U{L.rows * L.cols, W, widx, (int)lane, (int)warp};
        const int passes = U.passes();

        for(int it = 0; it <= 10; it++)
        {
            bool done;
            ++rel;
            if(is_solver_cta)
            {
                // ---- CTA 0: digest the previous evaluation (:348-380), publish the next homography ----
                if(it > 0) gather_rows(A.rows, W, kSo3Chunks, arr, s_rows, s_red, s_final);
                if(is_solver)
                {
                    int d = 0;
                    if(it > 0) d = solve_so3(S, s_final, it);
                    make_so3_params(S, s_par[rel & 1u], d, L.fx, L.fy, L.cx, L.cy, L.K_inv);
                    s_flag = d;
                }
                group_sync();
                if(warp == 0) warp_publish(A.par, s_par[rel & 1u], 0, kLineChunks, rel);
                done = s_flag != 0;
            }
            else
            {
                float * par = s_par[rel & 1u];
                wait_chunks(my_par, 0, kLineChunks, rel, par);
                done = par[18] != 0.f;
                P.image_basis = mat_from(par);
                P.krlr = mat_from(par + 9);
            }
            if(done) break;
            ++arr;
            if(is_solver_cta) continue;

            // ---- workers: so3Step over this CTA's pixels ----
            float acc[kSets][16];
#pragma unroll
            for(int q = 0; q < kSets; q++)
#pragma unroll
                for(int i = 0; i < 16; i++) acc[q][i] = 0.f;
            for(int p = 0; p < passes; p += kSets)
            {
#pragma unroll
                for(int q = 0; q < kSets; q++) // pass p + q belongs to virtual warp set q (p is a multiple of kSets)
                {
                    const int u = (p + q < passes) ? U.unit(p + q) : -1;
                    if(u >= 0)
                    {
                        const int y = u / L.cols, x = u - y * L.cols;
                        float row[4];
                        if(so3_row(P, x, y, A.so3_last, A.so3_next, L.cols, row)) accumulate_so3(acc[q], row);
                    }
                }
            }
#pragma unroll
            for(int q = 0; q < kSets; q++)
            {
                const float lane_value = warp_transpose_reduce16(acc[q]);
                if(lane < 16) s_red[(warp + q * kWarps) * 16 + lane] = lane_value;
            }
            group_sync();
            if(gtid() < 16)
            {
                float s = 0.f;
#pragma unroll
                for(int w = 0; w < kVWarps; w++) s += s_red[w * 16 + gtid()];
                s_final[gtid()] = (gtid() < 11) ? s : 0.f;
            }
            group_sync();
            publish_row(A.rows + (size_t)widx * kSo3Chunks, s_final, kSo3Chunks, arr);
        }
        if(is_solver)
        {
#pragma unroll
            for(int x = 0; x < 3; x++)
            {
#pragma unroll
                for(int y = 0; y < 3; y++) S->resultRt[x * 4 + y] = S->so3_R[x * 3 + y]; // :394-403
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
    auto solve_pending = [&](bool hand_over) {
        gather_rows(A.rows, W, kRowChunks, arr, s_rows, s_red, s_final);
        stamp(2);
        if(warp == 0)
        {
            warp_solve_se3(S, s_final, A.icp, A.rgb, A.icp_weight, pending_level, hand_over,
                           (TIMING && dbg_it < kMaxIters) ? A.dbg + (size_t)dbg_it * kDbgStamps : nullptr);
        }
        stamp(3);
    };

    // Level preparation in units (one compaction group or one staging batch each, ~5 000 cycles): a level's units run at its
    // start at the latest, but when all levels' lists fit shared memory side by side (A.prework) the workers run them earlier,
    // one per iteration of a coarser level, in the window where they would only wait for the next parameters.
    int prep_unit[kNumPyrs] = {0, 0, 0}, prep_cand[kNumPyrs] = {0, 0, 0};
    auto prep_passes = [&](int lv) { return UnitIter{A.lvl[lv].rows * A.lvl[lv].cols, W, widx, (int)lane, (int)warp}.passes(); };
    auto prep_groups = [&](int lv) { return A.rgb ? (prep_passes(lv) + kCompactGroup - 1) / kCompactGroup : 0; };
    auto prep_units = [&](int lv) {
        return prep_groups(lv) + ((A.icp && A.icp_in_smem) ? (prep_passes(lv) + kStageBatch - 1) / kStageBatch : 0);
    };
    auto prepare_unit = [&](int lv) { // all threads of a worker CTA
        const LevelArgs & PL = A.lvl[lv];
        const UnitIter PU{PL.rows * PL.cols, W, widx, (int)lane, (int)warp};
        const int pp = PU.passes(), groups = prep_groups(lv), unit = prep_unit[lv]++;
        if(unit < groups)
        {
            RgbResParams PR;
            PR.min_scale = PL.min_scale;
            PR.max_depth_delta = A.max_depth_delta;
            PR.rows = PL.rows;
            PR.cols = PL.cols;
            const CandStore PC = cand_store(lv);
            prep_cand[lv] = A.make_derivatives ? compact_group<true>(PL, PR, PU, pp, PC, s_wtot, unit, prep_cand[lv])
                                               : compact_group<false>(PL, PR, PU, pp, PC, s_wtot, unit, prep_cand[lv]);
        }
        else
            stage_batch(PL, PU, pp, vn_store(lv), A.lvl_cap[lv], (unit - groups) * kStageBatch);
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
        UnitIter U{L.rows * L.cols, W, widx, (int)lane, (int)warp};
        const int passes = U.passes();

        // ---- workers, level start: whatever is left of this level's preparation (candidate list, staged current maps) ----
        const CandStore C = cand_store(lv);
        float * const s_vn = vn_store(lv);
        dbg_it = it_global;
        if(!is_solver_cta) stamp(10);
        if(!is_solver_cta)
        {
            if(!A.prework) group_sync(); // the levels share one region: every thread is done with the previous level's records
            while(prep_unit[lv] < prep_units(lv)) prepare_unit(lv);
            group_sync();
        }
        if(!is_solver_cta) stamp(11);
        if(!is_solver_cta) stamp(12);
        const int n_cand = prep_cand[lv];
        // the next finer level with iterations: prepared one unit per iteration while this level's workers wait for parameters
        int next_lv = lv - 1;
        while(next_lv >= 0 && A.lvl[next_lv].iterations <= 0) next_lv--;

        float lastRGBError = FLT_MAX;      // every thread tracks it for the uniform rgb_only break (:464)
        bool first_of_level = true;

        for(int j = 0; j < L.iterations; j++)
        {
            const int cnt_slot = it_global++; // one barrier-B word per started iteration (also when it breaks)
            dbg_it = cnt_slot;
            stamp(0);
            ++rel;
            // ---- CTA 0: finish the previous iteration; publish this one's pose (:480-481) at once and the photometric
            //      warp (:424-434) when it is ready -- the workers run their first ICP pixels in between ----
            // ICP passes done before the photometric association (EF_TRACK_ICP_SPLIT: measured slower at 640x480 -- two
            // short batches cost more L2 round trips than the hidden parameter latency saves -- so off by default)
            const int split = (kIcpSplit && A.icp && A.rgb && passes >= 4) ? passes / 2 : 0;
            float accI[kSets][32];
#pragma unroll
            for(int q = 0; q < kSets; q++)
#pragma unroll
                for(int i = 0; i < 32; i++) accI[q][i] = 0.f;
            bool anyI = false;
            if(is_solver_cta)
            {
                const bool was_pending = pending;
                if(pending) solve_pending(A.rgb != 0); // warp 1 takes resultRt over right after the SE(3) update
                if(is_solver && first_of_level) S->last_rgb_error = FLT_MAX; // :420
                if(warp == 0)
                {
                    // (with a solve pending the hand-over to warp 1 happened inside warp_solve_se3, right after the SE(3) update)
                    if(!was_pending && A.rgb) solver_pair_sync();
                    float * par = s_par[rel & 1u];
                    warp_make_pose(S, par);
                    warp_publish(A.par, par, 0, 4, rel);
                    stamp(7);
                }
                else if(warp == 1 && A.rgb)
                {
                    // the photometric warp K R K^-1, K t of this iteration, in parallel with warp 0's pose composition
                    solver_pair_sync();
                    float * par = s_par[rel & 1u];
                    warp_make_rgb_params(S, par, L.fx, L.fy, L.cx, L.cy, L.K_inv);
                    warp_publish(A.par, par, 4, 4, rel);
                }
                stamp(1);
            }
            else
            {
                float * par = s_par[rel & 1u];
                if(split > 0)
                {
                    wait_chunks(my_par, 0, 4, rel, par);
                    IP.Rcurr = mat_from(par);
                    IP.tcurr = make_float3(par[9], par[10], par[11]);
                    stamp(4);
                    // ---- workers, phase A0: the first ICP pixels of every thread while CTA 0 computes the warp ----
                    anyI |= A.icp_in_smem ? icp_passes<true>(L, IP, U, 0, split, accI, s_vn, A.lvl_cap[lv]) : icp_passes<false>(L, IP, U, 0, split, accI, s_vn, A.lvl_cap[lv]);
                    wait_chunks(my_par, 4, 4, rel, par);
                }
                else
                {
                    // a unit of the next level's preparation in the window where this CTA would only wait (not before the first
                    // iteration of a level: its parameters are already there)
                    if(A.prework && j > 0)
                    {
                        while(next_lv >= 0 && prep_unit[next_lv] >= prep_units(next_lv))
                        {
                            next_lv--;
                            while(next_lv >= 0 && A.lvl[next_lv].iterations <= 0) next_lv--;
                        }
                        if(next_lv >= 0) prepare_unit(next_lv);
                    }
                    wait_chunks(my_par, A.icp ? 0 : 4, (A.icp && A.rgb) ? 8 : 4, rel, par);
                    IP.Rcurr = mat_from(par);
                    IP.tcurr = make_float3(par[9], par[10], par[11]);
                    stamp(4);
                }
                RP.krkinv = mat_from(par + 12);
                RP.kt = make_float3(par[21], par[22], par[23]);
            }
            pending = false;
            first_of_level = false;

            // ---- workers, phase A1: photometric association of the candidates -> records in shared memory, then the
            //      barrier-B arrival: this CTA's {count, sum diff^2} as one flagged chunk in its own slot (integer sums are
            //      exact in any order and sigma wraps mod 2^32 exactly like the reference's int sum) ----
            if(!is_solver_cta && A.rgb)
            {
                int cnt = 0, sig = 0;
                rgb_assoc_cands(L, RP, C, n_cand, cnt, sig);
                cnt = __reduce_add_sync(kFullMask, cnt);
                sig = __reduce_add_sync(kFullMask, sig);
                if(lane == 0) { s_wcnt[warp] = cnt; s_wsig[warp] = sig; }
                group_sync();
                if(gtid() == 0)
                {
                    unsigned c = 0, s = 0;
#pragma unroll
                    for(int w = 0; w < kWarps; w++) { c += (unsigned)s_wcnt[w]; s += (unsigned)s_wsig[w]; }
                    st_relaxed_v4(A.bslot + widx, make_uint4(c, s, 0u, rel));
                }
            }
            if(!is_solver_cta) stamp(5);

            // ---- workers, phase A2: the remaining ICP pixels (hides the barrier-B round trip and the inter-CTA skew) ----
            if(!is_solver_cta)
            {
                float vi[kSets];
#pragma unroll
                for(int q = 0; q < kSets; q++) vi[q] = 0.f;
                if(A.icp)
                {
                    anyI |= A.icp_in_smem ? icp_passes<true>(L, IP, U, split, passes, accI, s_vn, A.lvl_cap[lv])
                                          : icp_passes<false>(L, IP, U, split, passes, accI, s_vn, A.lvl_cap[lv]);
                    if(__any_sync(kFullMask, anyI)) // a warp without pixels contributes zeros
                    {
#pragma unroll
                        for(int q = 0; q < kSets; q++) vi[q] = warp_transpose_reduce32(accI[q]);
                    }
                }
#pragma unroll
                for(int q = 0; q < kSets; q++)
                    if(lane < 29) s_red[(warp + q * kWarps) * 64 + lane] = vi[q];
                stamp(8);
            }

            // ---- barrier B: CTA 0 (idle while the workers compute) collects the arrivals, one per thread, adds them and
            //      hands every worker the robust-weight scale (:461-462) in its own inbox; the workers get there after their
            //      ICP phase ----
            bool level_break = false;
            if(A.rgb)
            {
                if(is_solver_cta)
                {
                    int c = 0, sg = 0;
                    for(int w = gtid(); w < W; w += kThreads)
                    {
                        uint4 v;
                        do
                        {
                            v = ld_relaxed_v4(A.bslot + w);
                        } while(v.w != rel);
                        c += (int)v.x;
                        sg += (int)v.y;
                    }
                    c = __reduce_add_sync(kFullMask, c);
                    sg = __reduce_add_sync(kFullMask, sg);
                    if(lane == 0) { s_wcnt[warp] = c; s_wsig[warp] = sg; }
                    group_sync();
                    if(gtid() == 0)
                    {
                        unsigned cc = 0, ss = 0;
#pragma unroll
                        for(int w = 0; w < kWarps; w++) { cc += (unsigned)s_wcnt[w]; ss += (unsigned)s_wsig[w]; }
                        float sigmaVal, rgbError;
                        sigma_from_sums((int)ss, (int)cc, sigmaVal, rgbError); // :461-462
                        s_sigma[0] = sigmaVal;
                        s_sigma[1] = rgbError;
                        s_sigma[2] = (float)(int)cc;
                    }
                    group_sync();
                    for(int w = gtid(); w < W; w += kThreads)
                        st_relaxed_v4(A.bres + w, make_uint4(__float_as_uint(s_sigma[0]), __float_as_uint(s_sigma[1]), __float_as_uint(s_sigma[2]), rel));
                }
                else
                {
                    if(gtid() == 0)
                    {
                        uint4 v;
                        do
                        {
                            v = ld_relaxed_v4(A.bres + widx);
                        } while(v.w != rel);
                        s_sigma[0] = __uint_as_float(v.x);
                        s_sigma[1] = __uint_as_float(v.y);
                        s_sigma[2] = __uint_as_float(v.z);
                    }
                    group_sync();
                }
                stamp(6);
                const float rgbError = s_sigma[1];
                if(A.rgb_only && rgbError > lastRGBError) level_break = true; // :464 (uniform across the grid)
                if(!level_break)
                {
                    lastRGBError = rgbError;
                    if(is_solver)
                    {
                        S->last_rgb_error = rgbError; // :469-470
                        S->last_rgb_count = s_sigma[2];
                    }
                    SP.sigma = A.rgb_only ? -1.f : s_sigma[0]; // :472-475
                }
            }
            else if(is_solver)
            {
                // :461-470 run even without RGB: sigma = rgbSize = 0 -> rgbError 0, count 0
                S->last_rgb_error = 0.f;
                S->last_rgb_count = 0.f;
            }
            if(level_break) break; // no row outstanding: every CTA takes the same branch

            ++arr;
            pending = true;
            pending_level = lv;
            if(is_solver_cta) continue;

            // ---- workers, phase B: photometric rows from the records in shared memory -> 29 more sums ----
            {
                float vr[kSets];
#pragma unroll
                for(int q = 0; q < kSets; q++) vr[q] = 0.f;
                if(A.rgb)
                {
                    float accR[kSets][32];
#pragma unroll
                    for(int q = 0; q < kSets; q++)
#pragma unroll
                        for(int i = 0; i < 32; i++) accR[q][i] = 0.f;
                    const bool any = rgb_rows_cands(SP, C, n_cand, accR);
                    if(__any_sync(kFullMask, any))
                    {
#pragma unroll
                        for(int q = 0; q < kSets; q++) vr[q] = warp_transpose_reduce32(accR[q]);
                    }
                }
#pragma unroll
                for(int q = 0; q < kSets; q++)
                    if(lane < 29) s_red[(warp + q * kWarps) * 64 + 29 + lane] = vr[q];
                group_sync();
                if(gtid() < kRowFloats)
                {
                    float sum = 0.f;
                    if(gtid() < 58)
                    {
#pragma unroll
                        for(int w = 0; w < kVWarps; w++) sum += s_red[w * 64 + gtid()];
                    }
                    s_final[gtid()] = sum;
                }
                group_sync();
                publish_row(A.rows + (size_t)widx * kRowChunks, s_final, kRowChunks, arr);
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
            group_sync();
            publish_row(A.rows + (size_t)widx * kSo3Chunks, s_final, kSo3Chunks, arr);
        }
    }
    if(is_solver_cta)
    {
        if(pending) solve_pending(false);
        else gather_rows(A.rows, W, kSo3Chunks, arr, s_rows, s_red, s_final);
    }
    if(is_solver)
    {
        float Rc[9], tc[3];
#pragma unroll
        for(int i = 0; i < 9; i++) Rc[i] = S->Rcurr[i];
#pragma unroll
        for(int i = 0; i < 3; i++) tc[i] = S->tcurr[i];
        if(A.rgb)
        {
            const float d[3] = {tc[0] - A.tprev[0], tc[1] - A.tprev[1], tc[2] - A.tprev[2]};
            if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
            {
#pragma unroll
                for(int i = 0; i < 9; i++) Rc[i] = A.Rprev[i];
#pragma unroll
                for(int i = 0; i < 3; i++) tc[i] = A.tprev[i];
            }
        }
        TrackOutput * out = A.out;
#pragma unroll
        for(int i = 0; i < 3; i++) out->trans[i] = tc[i];
#pragma unroll
        for(int i = 0; i < 9; i++) out->rot[i] = Rc[i];
        out->st.last_icp_error = S->last_icp_error; out->st.last_icp_count = S->last_icp_count;
        out->st.last_rgb_error = S->last_rgb_error; out->st.last_rgb_count = S->last_rgb_count;
        out->st.last_so3_error = S->last_so3_error; out->st.last_so3_count = S->last_so3_count;
        out->st.so3_iterations = S->so3_iterations;
#pragma unroll
        for(int i = 0; i < 3; i++) out->st.se3_iterations[i] = S->se3_iterations[i];
        if(S->se3_iterations[0] + S->se3_iterations[1] + S->se3_iterations[2] > 0)
        {
            // lastA / lastb of the last solve (reduce.cu:475-486 unpack order)
#pragma unroll
            for(int i = 0; i < 6; i++)
            {
#pragma unroll
                for(int j = i; j < 7; j++)
                {
                    const double v = S->last_S[hm::acc_index(i, j)];
                    if(j == 6) out->st.last_b[i] = v;
                    else out->st.last_A[j * 6 + i] = out->st.last_A[i * 6 + j] = v;
                }
            }
        }
        // the host polls `status` in pinned memory (device_track_finish): result first, fence, then the flag.
        // 2 = no solve ran: lastA / lastb keep their previous values (host side)
        const int status = (S->se3_iterations[0] + S->se3_iterations[1] + S->se3_iterations[2] > 0) ? 1 : 2;
        if constexpr(TIMING) A.dbg[(size_t)(kMaxIters - 1) * kDbgStamps] = clock64();
        __threadfence_system();
        *reinterpret_cast<volatile int *>(&out->status) = status;
        __threadfence_system();
    }
}

// CTA 0 gathers and solves, CTAs 1 .. grid-1 are the workers (of every thread group's sequence in the batched build)
template<bool TIMING>
__global__ void __launch_bounds__(kThreads * kGroups, 1) k_track(const __grid_constant__ BatchArgs BA)
{
    extern __shared__ int4 s_dyn_all[];
    __shared__ GroupShared s_groups[kGroups];
    const TrackArgs & A = BA.seq[ggrp()];
    track_body<TIMING>(A, s_groups[ggrp()], s_dyn_all + (size_t)ggrp() * (A.group_smem_bytes / sizeof(int4)), blockIdx.x == 0,
                       blockIdx.x == 0 ? 0 : (int)blockIdx.x - 1, (int)gridDim.x - 1);
}

#if defined(EF_TRACK_ALT)
// ------------------------------------------------------------------------------------------------
// The alternating build: k sequences per launch, the worker CTAs time-sliced between them
// ------------------------------------------------------------------------------------------------
constexpr int kAltMax = 4; // sequences per launch (each takes one solver CTA and a share of the workers' shared memory)

struct AltArgs
{
    TrackArgs seq[kAltMax];
    int n;
};

// where a worker CTA stands in one of its sequences (shared memory; identical in every worker CTA of the grid)
struct WorkerSeq
{
    unsigned rel, arr;   // the epochs of track_body
    int stage;           // 0 SO(3) pre-alignment | 1 Gauss-Newton | 2 finished
    int lv, j;           // stage 1: the iteration to run next
    int pending;         // a row round the solver has not digested yet
    int n_cand;          // photometric candidates of this CTA at level lv
    float lastRGBError;
};

__device__ __forceinline__ CandStore alt_cand_store(const TrackArgs & A, int4 * s_dyn, int lv)
{
    CandStore c;
    unsigned * base = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(s_dyn) + A.lvl_off[lv]);
    const int cap = A.lvl_cap[lv];
    c.c0 = base;
    c.c1 = base + cap;
    c.c2 = reinterpret_cast<float *>(base + 2 * cap);
    c.r0 = base + 3 * cap;
    c.r1 = reinterpret_cast<float *>(base + 4 * cap);
    return c;
}

// level lv of a sequence has iterations left to start: the first such level at or below `lv`, -1 if none
__device__ __forceinline__ int alt_next_level(const TrackArgs & A, int lv)
{
    while(lv >= 0 && A.lvl[lv].iterations <= 0) lv--;
    return lv;
}

// ONE step of one sequence in a worker CTA: an so3Step evaluation or a Gauss-Newton iteration -- from the wait for its
// parameters (published long ago when the other sequences kept this CTA busy meanwhile) to the publication of this CTA's
// partial sums.  The worker half of track_body, statement for statement (same pixels, same order of additions: a sequence
// gets the bits of a single launch with the same number of workers); the loop state lives in `w`.
__device__ __forceinline__ void alt_worker_step(const TrackArgs & A, GroupShared & GS, int4 * const s_dyn, WorkerSeq & w, const int widx, const int W)
{
    float * const s_red = GS.red;
    float * const s_final = GS.final_;
    float (* const s_par)[kPayload] = GS.par;
    float * const s_sigma = GS.sigma;
    int * const s_wcnt = GS.wcnt, * const s_wsig = GS.wsig;
    const uint4 * my_par = A.par + (size_t)(widx % kReplicas) * kReplicaStride;
    const unsigned lane = gtid() & 31u, warp = gtid() >> 5;

    if(w.stage == 0)
    {
        // ---- SO(3) pre-alignment (RGBDOdometry.cpp:294-382): one so3Step evaluation, or the news that it is over ----
        const LevelArgs & L = A.lvl[2];
        ++w.rel;
        float * par = s_par[w.rel & 1u];
        wait_chunks(my_par, 0, kLineChunks, w.rel, par);
        if(par[18] != 0.f)
        {
            w.stage = 1;
            return;
        }
        So3Params P;
        P.rows = L.rows;
        P.cols = L.cols;
        P.kinv = mat_from(A.so3_kinv);
        P.image_basis = mat_from(par);
        P.krlr = mat_from(par + 9);
        ++w.arr;
        const UnitIter U{L.rows * L.cols, W, widx, (int)lane, (int)warp};
        const int passes = U.passes();
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
        group_sync();
        if(gtid() < 16)
        {
            float sum = 0.f;
#pragma unroll
            for(int k = 0; k < kWarps; k++) sum += s_red[k * 16 + gtid()];
            s_final[gtid()] = (gtid() < 11) ? sum : 0.f;
        }
        group_sync();
        publish_row(A.rows + (size_t)widx * kSo3Chunks, s_final, kSo3Chunks, w.arr);
        return;
    }

    // ---- one Gauss-Newton iteration (RGBDOdometry.cpp:405-585) of level w.lv ----
    const int lv = w.lv;
    const LevelArgs & L = A.lvl[lv];
    IcpParams IP;
    IP.Rprev_inv = mat_from(A.Rprev_inv);
    IP.tprev = make_float3(A.tprev[0], A.tprev[1], A.tprev[2]);
    IP.dist_thresh = A.dist_thresh;
    IP.angle_thresh = A.angle_thresh;
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
    const UnitIter U{L.rows * L.cols, W, widx, (int)lane, (int)warp};
    const int passes = U.passes();
    const CandStore C = alt_cand_store(A, s_dyn, lv);
    float * const s_vn = reinterpret_cast<float *>(reinterpret_cast<char *>(s_dyn) + A.lvl_off[lv]) + 5 * (size_t)A.lvl_cap[lv];

    if(w.j == 0)
    {
        // level start: candidate list and staged current maps of this CTA's pixels (the levels of a sequence share one region;
        // every thread is past the previous level's records: the previous step of this sequence ended with CTA barriers)
        int n_cand = 0;
        if(A.rgb)
        {
            const int groups = (passes + kCompactGroup - 1) / kCompactGroup;
            for(int g = 0; g < groups; g++)
                n_cand = A.make_derivatives ? compact_group<true>(L, RP, U, passes, C, GS.wtot, g, n_cand) : compact_group<false>(L, RP, U, passes, C, GS.wtot, g, n_cand);
        }
        if(A.icp && A.icp_in_smem)
            for(int p0 = 0; p0 < passes; p0 += kStageBatch) stage_batch(L, U, passes, s_vn, A.lvl_cap[lv], p0);
        group_sync();
        w.n_cand = n_cand;
        w.lastRGBError = FLT_MAX;
    }
    const int n_cand = w.n_cand;

    ++w.rel;
    float * par = s_par[w.rel & 1u];
    wait_chunks(my_par, A.icp ? 0 : 4, (A.icp && A.rgb) ? 8 : 4, w.rel, par);
    IP.Rcurr = mat_from(par);
    IP.tcurr = make_float3(par[9], par[10], par[11]);
    RP.krkinv = mat_from(par + 12);
    RP.kt = make_float3(par[21], par[22], par[23]);
    w.pending = 0;

    // phase A1: photometric association -> records in shared memory; this CTA's barrier-B arrival
    if(A.rgb)
    {
        int cnt = 0, sig = 0;
        rgb_assoc_cands(L, RP, C, n_cand, cnt, sig);
        cnt = __reduce_add_sync(kFullMask, cnt);
        sig = __reduce_add_sync(kFullMask, sig);
        if(lane == 0) { s_wcnt[warp] = cnt; s_wsig[warp] = sig; }
        group_sync();
        if(gtid() == 0)
        {
            unsigned c = 0, sg = 0;
#pragma unroll
            for(int k = 0; k < kWarps; k++) { c += (unsigned)s_wcnt[k]; sg += (unsigned)s_wsig[k]; }
            st_relaxed_v4(A.bslot + widx, make_uint4(c, sg, 0u, w.rel));
        }
    }

    // phase A2: ICP association + 29 sums (hides the barrier-B round trip)
    {
        float accI[1][32];
#pragma unroll
        for(int i = 0; i < 32; i++) accI[0][i] = 0.f;
        float vi = 0.f;
        if(A.icp)
        {
            const bool anyI = A.icp_in_smem ? icp_passes<true>(L, IP, U, 0, passes, accI, s_vn, A.lvl_cap[lv]) : icp_passes<false>(L, IP, U, 0, passes, accI, s_vn, A.lvl_cap[lv]);
            if(__any_sync(kFullMask, anyI)) vi = warp_transpose_reduce32(accI[0]);
        }
        if(lane < 29) s_red[warp * 64 + lane] = vi;
    }

    // barrier B: the robust-weight scale from the sequence's solver CTA
    bool level_break = false;
    if(A.rgb)
    {
        if(gtid() == 0)
        {
            uint4 v;
            do
            {
                v = ld_relaxed_v4(A.bres + widx);
            } while(v.w != w.rel);
            s_sigma[0] = __uint_as_float(v.x);
            s_sigma[1] = __uint_as_float(v.y);
            s_sigma[2] = __uint_as_float(v.z);
        }
        group_sync();
        const float rgbError = s_sigma[1];
        if(A.rgb_only && rgbError > w.lastRGBError) level_break = true; // :464
        if(!level_break)
        {
            w.lastRGBError = rgbError;
            SP.sigma = A.rgb_only ? -1.f : s_sigma[0]; // :472-475
        }
    }

    if(!level_break)
    {
        ++w.arr;
        w.pending = 1;
        // phase B: photometric rows from the records -> 29 more sums; the CTA's row of this round
        float vr = 0.f;
        if(A.rgb)
        {
            float accR[1][32];
#pragma unroll
            for(int i = 0; i < 32; i++) accR[0][i] = 0.f;
            const bool any = rgb_rows_cands(SP, C, n_cand, accR);
            if(__any_sync(kFullMask, any)) vr = warp_transpose_reduce32(accR[0]);
        }
        if(lane < 29) s_red[warp * 64 + 29 + lane] = vr;
        group_sync();
        if(gtid() < kRowFloats)
        {
            float sum = 0.f;
            if(gtid() < 58)
            {
#pragma unroll
                for(int k = 0; k < kWarps; k++) sum += s_red[k * 64 + gtid()];
            }
            s_final[gtid()] = sum;
        }
        group_sync();
        publish_row(A.rows + (size_t)widx * kRowChunks, s_final, kRowChunks, w.arr);
        w.j++;
    }
    if(level_break || w.j >= L.iterations)
    {
        w.j = 0;
        w.lv = alt_next_level(A, lv - 1);
        if(w.lv < 0)
        {
            // epilogue: without a round outstanding, a round without payload tells the solver that every worker is past its
            // last barrier-B wait
            if(!w.pending)
            {
                ++w.arr;
                group_sync();
                publish_row(A.rows + (size_t)widx * kSo3Chunks, s_final, kSo3Chunks, w.arr);
            }
            w.stage = 2;
        }
    }
}

// Blocks 0 .. n-1: the solver CTAs of the n sequences (track_body's solver half, unchanged).  Blocks n .. grid-1: the workers,
// which take the sequences in turn, one step each.  A worker never waits for anything that depends on its own future steps
// (sequence s's parameters need only the rows of step-1 of sequence s from every worker), so the rotation cannot deadlock.
__global__ void __launch_bounds__(kThreads, 1) k_track_alt(const __grid_constant__ AltArgs BA)
{
    extern __shared__ int4 s_dyn_all[];
    __shared__ GroupShared s_group;
    __shared__ WorkerSeq s_ws[kAltMax];
    const int n = BA.n;
    const int W = (int)gridDim.x - n;
    if((int)blockIdx.x < n)
    {
        track_body<false>(BA.seq[blockIdx.x], s_group, s_dyn_all, true, 0, W);
        return;
    }
    const int widx = (int)blockIdx.x - n;
    if((int)threadIdx.x < n)
    {
        const TrackArgs & A = BA.seq[threadIdx.x];
        WorkerSeq w;
        w.rel = w.arr = A.epoch_base;
        w.lv = alt_next_level(A, kNumPyrs - 1);
        w.j = 0;
        w.stage = A.so3 ? 0 : 1;
        w.pending = 0;
        w.n_cand = 0;
        w.lastRGBError = FLT_MAX;
        s_ws[threadIdx.x] = w;
    }
    __syncthreads();
    for(int left = n; left > 0;)
    {
        for(int s = 0; s < n; s++)
        {
            WorkerSeq w = s_ws[s];
            if(w.stage == 2) continue;
            alt_worker_step(BA.seq[s], s_group, s_dyn_all + (size_t)s * (BA.seq[s].group_smem_bytes / sizeof(int4)), w, widx, W);
            if(w.stage == 2) left--;
            __syncthreads(); // every thread has its copy of s_ws[s] (and is done with the step's shared memory)
            if(threadIdx.x == 0) s_ws[s] = w;
            __syncthreads();
        }
    }
}
#endif // EF_TRACK_ALT

struct Layout // shared-memory geometry of one thread group for a grid of `grid` CTAs
{
    int grid;
    int cand_cap, icp_in_smem, prework;
    int lvl_cap[kNumPyrs], lvl_off[kNumPyrs];
    int cand_base;     // bytes: where the lists start in the group's dynamic shared memory (symmetric body: behind the gathered rows)
    size_t smem_bytes; // per group, multiple of 16
};

struct DeviceTrack
{
    long long * dbg;      // device, max_grid * kMaxIters * kDbgStamps stamps (EF_TRACK_TIMING=1)
    double dbg_acc[kMaxIters][kDbgStamps];
    double wrk_mean[kMaxIters][7], wrk_max[kMaxIters][7]; // worker phase durations, mean / max over the worker CTAs
    long long * dbg_host;
    int dbg_grid;
    long long dbg_n;
    uint4 * par, * bslot, * bres;
    uint4 * rows;
    TrackOutput * out; // pinned
    Layout lay;        // of the single launch (all of a CTA's shared memory)
    Layout lay_sym;    // the same grid with the symmetric body
    bool use_sym;      // this handle's launches run k_track_sym (a handle on a subset of the SMs)
    unsigned launch_seq;
};

// pixels per thread unit and passes per level -> shared-memory records per thread, within `budget` bytes of dynamic
// shared memory.  false: the photometric candidates of a CTA do not fit.
// solvers = 1: track_body (CTA 0 gathers and solves); n: the alternating build; 0: the symmetric body -- every CTA is a worker and
// gathers all rows itself (rows and lists side by side, one level's lists at a time)
bool compute_layout(const ef_tracker * t, int grid, size_t budget, Layout & L, int solvers = 1, bool allow_prework = true)
{
    const int W = grid - solvers;
    const size_t rows_beside = solvers ? 0 : (((size_t)grid * kRowFloats * sizeof(float) + 15) & ~(size_t)15);
    L.cand_base = (int)rows_beside;
    if(!solvers)
    {
        if(budget <= rows_beside) return false;
        budget -= rows_beside;
        allow_prework = false;
    }
    int max_cand = 1;
    for(int i = 0; i < kNumPyrs; i++)
    {
        const int npix = t->dims[i].rows * t->dims[i].cols;
        const int chunks = (npix + 31) / 32;
        const int per_worker = (chunks + W - 1) / W;
        if(per_worker * 32 > max_cand) max_cand = per_worker * 32;
    }
    size_t smem = (size_t)max_cand * kCandBytes;
    const bool icp_in_smem = (size_t)max_cand * (kCandBytes + kIcpBytes) <= budget;
    if(icp_in_smem) smem = (size_t)max_cand * (kCandBytes + kIcpBytes);
    // the lists of the three levels side by side, if they fit: finer levels are then prepared while coarser ones iterate
    const size_t per_px = kCandBytes + (icp_in_smem ? kIcpBytes : 0);
    size_t all = 0;
    for(int i = 0; i < kNumPyrs; i++)
    {
        const int npix = t->dims[i].rows * t->dims[i].cols;
        const int per_worker = ((npix + 31) / 32 + W - 1) / W;
        L.lvl_cap[i] = per_worker * 32;
        L.lvl_off[i] = (int)all;
        all += ((size_t)L.lvl_cap[i] * per_px + 15) & ~(size_t)15;
    }
    static const bool no_prework_env = getenv("EF_TRACK_PREWORK") && getenv("EF_TRACK_PREWORK")[0] == '0'; // (experiments)
    const bool prework = allow_prework && !no_prework_env && all <= budget;
    if(prework) smem = all;
    else
        for(int i = 0; i < kNumPyrs; i++)
        {
            L.lvl_cap[i] = max_cand; // one region, reused level after level
            L.lvl_off[i] = 0;
        }
    const size_t rows_smem = (size_t)W * kRowFloats * sizeof(float);
    if(solvers && rows_smem > smem) smem = rows_smem;
    smem += rows_beside;
    L.grid = grid;
    L.prework = prework ? 1 : 0;
    L.cand_cap = max_cand;
    L.icp_in_smem = icp_in_smem ? 1 : 0;
    L.smem_bytes = (smem + 15) & ~(size_t)15;
    return L.smem_bytes <= budget + rows_beside;
}

// everything a launch passes to one thread group (one sequence)
void fill_args(ef_tracker * t, DeviceTrack * d, const Layout & lay, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid,
               int fast_odom, int so3, TrackArgs & A)
{
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
    // launch-unique flag epochs: nothing in the control block or the rows needs resetting between launches.  After
    // 2^24 launches the sequence wraps; stale flags are wiped then.
    d->launch_seq++;
    if((d->launch_seq & 0xffffffu) == 0)
    {
        d->launch_seq = 1;
        cudaMemsetAsync(d->rows, 0, kRowsChunks * sizeof(uint4), t->stream);
        cudaMemsetAsync(d->par, 0, kCtlChunks * sizeof(uint4), t->stream);
    }
    A.epoch_base = d->launch_seq << 8;
    A.par = d->par;
    A.bslot = d->bslot;
    A.bres = d->bres;
    A.rows = d->rows;
    A.out = d->out; // UVA: pinned + mapped host memory is addressable from the device
    A.dbg = d->dbg;
    A.cand_cap = lay.cand_cap;
    A.prework = lay.prework;
    for(int i = 0; i < kNumPyrs; i++)
    {
        A.lvl_cap[i] = lay.lvl_cap[i];
        A.lvl_off[i] = lay.lvl_off[i];
    }
    A.make_derivatives = (A.rgb && !t->deriv_valid) ? 1 : 0;
    A.icp_in_smem = lay.icp_in_smem;
    A.group_smem_bytes = (int)lay.smem_bytes;
    A.cand_base = lay.cand_base;
    d->out->status = 0;
}

} // namespace

#if defined(EF_TRACK_ALT)
// n sequences (handles of one image size on one device, pyramids built, streams joined) from ONE launch on `stream`.  Every
// sequence gets an n-th of the workers' shared memory for one level's lists at a time.  EF_ERR_UNSUPPORTED when that share
// cannot hold them.  grid_ctas <= 0: every SM.
int EF_TRACK_FN(device_track_launch_batch)(ef_tracker * const * ts, int n, const float * const * trans, const float * const * rot, int rgb_only,
                                           float icp_weight, int pyramid, int fast_odom, int so3, int grid_ctas, cudaStream_t stream)
{
    static bool attr_set[64] = {false}; // per device: function attributes belong to the device's context
    const int dev = ts[0]->device;
    if(dev < 0 || dev >= 64 || !attr_set[dev])
    {
        cudaError_t e = cudaFuncSetAttribute(k_track_alt, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if(e != cudaSuccess) return (int)e;
        if(dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if(n < 1 || n > kAltMax) return EF_ERR_INVALID_ARGUMENT;
    ef_tracker * t0 = ts[0];
    const int max_grid = t0->num_sms < kMaxGrid ? t0->num_sms : kMaxGrid;
    int grid = (grid_ctas > 0 && grid_ctas < max_grid) ? grid_ctas : max_grid;
    if(grid < n + 1) grid = n + 1;
    Layout L;
    if(grid > max_grid || !compute_layout(t0, grid, ((size_t)kMaxDynSmem / n) & ~(size_t)15, L, n, false))
    {
        t0->err = "batched launch: image too large for a sequence's share of the shared memory";
        return EF_ERR_UNSUPPORTED;
    }
    AltArgs local; // (~3 KB)
    AltArgs * Bp = &local;
    Bp->n = n;
    for(int g = 0; g < n; g++)
    {
        ef_tracker * t = ts[g];
        DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
        if(!d || t->width != t0->width || t->height != t0->height || t->device != t0->device) return EF_ERR_INVALID_ARGUMENT;
        fill_args(t, d, L, trans[g], rot[g], rgb_only, icp_weight, pyramid, fast_odom, so3, Bp->seq[g]);
        Bp->seq[g].dbg = nullptr; // the clock64 trace is a single-launch diagnostic
    }
    void * args[] = {Bp};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_track_alt, dim3(L.grid), dim3(kThreads), args, L.smem_bytes * n, stream);
    for(int g = 0; g < n; g++) ts[g]->launches++;
    if(e != cudaSuccess)
    {
        t0->err = std::string("cudaLaunchCooperativeKernel (alternating): ") + cudaGetErrorString(e);
        return (int)e;
    }
    return EF_OK;
}
#elif EF_TRACK_GROUPS == 1
// (re)derive the launch geometry for a grid of `grid` CTAs.  A handle normally owns every SM; EF_OPT_GRID_CTAS lets several
// handles share the GPU (e.g. two sequences tracked concurrently on 74 SMs each).
int EF_TRACK_FN(device_track_configure)(ef_tracker * t, int grid)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    const int max_grid = t->num_sms < kMaxGrid ? t->num_sms : kMaxGrid;
    if(grid <= 0 || grid > max_grid) grid = max_grid;
    if(grid < 2) grid = 2;
    Layout L;
    if(!compute_layout(t, grid, (size_t)kMaxDynSmem, L) && d->lay.grid > 0)
    {
        // too few CTAs for this image: the photometric candidates of a CTA must fit its shared memory; keep the old grid
        t->err = "EF_OPT_GRID_CTAS: too few CTAs for the shared-memory candidate store at this image size";
        return EF_ERR_UNSUPPORTED;
    }
    d->lay = L;
    // a handle that shares the GPU runs the symmetric body (see track_body_sym); EF_TRACK_SYM=0 / 1 forces one or the other
    const char * env = getenv("EF_TRACK_SYM");
    const bool want_sym = env ? env[0] == '1' : grid < max_grid;
    d->use_sym = want_sym && compute_layout(t, grid, (size_t)kMaxDynSmem, d->lay_sym, 0);
    return EF_OK;
}

bool EF_TRACK_FN(device_track_supported)(const ef_tracker * t)
{
    const DeviceTrack * d = static_cast<const DeviceTrack *>(t->track_state);
    const Layout & L = d && d->use_sym ? d->lay_sym : d->lay;
    return d && L.smem_bytes <= (size_t)kMaxDynSmem && (size_t)t->width * t->height < (1u << 24) && t->width <= 4094 && t->height <= 4094;
}

int EF_TRACK_FN(device_track_init)(ef_tracker * t)
{
    DeviceTrack * d = new DeviceTrack();
    memset(d, 0, sizeof(*d));
    t->track_state = d;
    EF_TRACK_FN(device_track_configure)(t, 0);
    const int max_grid = t->num_sms < kMaxGrid ? t->num_sms : kMaxGrid;
    d->launch_seq = 0;
    const size_t ctl_chunks = kCtlChunks;
    cudaError_t e = cudaMalloc((void **)&d->par, ctl_chunks * sizeof(uint4));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->par, 0, ctl_chunks * sizeof(uint4), t->stream);
    d->bslot = d->par ? d->par + 2 * (size_t)kParCopyStride : nullptr;
    d->bres = d->par ? d->bslot + 2 * kSlotCopy : nullptr;
    if(e == cudaSuccess) e = cudaMalloc((void **)&d->rows, kRowsChunks * sizeof(uint4));
    if(e == cudaSuccess) e = cudaMemsetAsync(d->rows, 0, kRowsChunks * sizeof(uint4), t->stream);
    if(e == cudaSuccess) e = cudaHostAlloc((void **)&d->out, sizeof(TrackOutput), cudaHostAllocMapped);
    const char * env = getenv("EF_TRACK_TIMING");
    if(e == cudaSuccess && env && env[0] == '1')
    {
        d->dbg_grid = max_grid;
        d->dbg_host = (long long *)malloc(sizeof(long long) * max_grid * kMaxIters * kDbgStamps);
        e = cudaMalloc((void **)&d->dbg, sizeof(long long) * max_grid * kMaxIters * kDbgStamps);
        if(e == cudaSuccess) e = cudaMemsetAsync(d->dbg, 0, sizeof(long long) * max_grid * kMaxIters * kDbgStamps, t->stream);
    }
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_track<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_track<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_track_sym<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_track_sym<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if(e != cudaSuccess)
    {
        EF_TRACK_FN(device_track_destroy)(t);
        return (int)e;
    }
    t->h_track_out = d->out;
    return EF_OK;
}

// EF_TRACK_TIMING=1: where CTA 0's clock stood at the top of every SE3 iteration, relative to the kernel's start, summed
// over the calls so far: out[i] for iteration i (0 where the iteration did not run), out[kMaxIters - 1] = kernel end
int EF_TRACK_FN(device_track_trace)(ef_tracker * t, double * out32, long long * calls)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d || !d->dbg) return EF_ERR_BAD_STATE;
    const double start = d->dbg_acc[kMaxIters - 2][0];
    for(int it = 0; it < kMaxIters; it++) out32[it] = d->dbg_acc[it][0] > 0 ? d->dbg_acc[it][0] - start : 0.0;
    out32[kMaxIters - 2] = 0.0;
    *calls = d->dbg_n;
    return EF_OK;
}

void EF_TRACK_FN(device_track_destroy)(ef_tracker * t)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return;
    if(d->dbg)
    {
        if(d->dbg_n > 0 && getenv("EF_TRACK_TIMING_PRINT"))
        {
            // stamps (cycles of CTA 0's SM): 0 loop top | 2 rows gathered + added | 3 solved | 1 parameters published | 6 barrier B seen
            fprintf(stderr, "[ef_track timing] avg cycles of CTA 0 per SE3 iteration over %lld calls\n", d->dbg_n);
            for(int it = 0; it < kMaxIters - 2; it++)
            {
                const double * a = d->dbg_acc[it];
                if(a[1] == 0) continue;
                const double n = (double)d->dbg_n;
                const bool first = a[2] == 0;
                fprintf(stderr, "  it %2d: gather %7.0f solve %7.0f (ldlt %5.0f backsub %5.0f exp %5.0f update+compose %5.0f) params %7.0f publish %7.0f | to-barB-seen %7.0f | total %8.0f\n", it,
                        first ? 0.0 : (a[2] - a[0]) / n, first ? 0.0 : (a[3] - a[2]) / n, a[8] / n, a[9] / n, a[5] / n, a[4] / n,
                        (a[7] - (first ? a[0] : a[3])) / n, (a[1] - a[7]) / n,
                        (a[6] - a[1]) / n, (a[6] - a[0]) / n);
            }
            // worker CTAs: params wait (0->4) | photometric association + CTA sync (4->5) | ICP + warp reduce (5->8) |
            // barrier-B wait (8->6) | photometric rows + CTA reduce + publish (6->9)
            fprintf(stderr, "[ef_track timing] worker CTAs, cycles mean/max over workers\n");
            for(int it = 0; it < kMaxIters - 2; it++)
            {
                if(d->dbg_acc[it][1] == 0) continue;
                const double n = (double)d->dbg_n;
                const double * m = d->wrk_mean[it];
                const double * x = d->wrk_max[it];
                fprintf(stderr, "  it %2d: wait-params %6.0f/%6.0f  rgb-assoc %6.0f/%6.0f  icp %6.0f/%6.0f  wait-barB %6.0f/%6.0f  rgb-rows+publish %6.0f/%6.0f", it,
                        m[0] / n, x[0] / n, m[1] / n, x[1] / n, m[2] / n, x[2] / n, m[3] / n, x[3] / n, m[4] / n, x[4] / n);
                if(m[5] > 0) fprintf(stderr, "  | level start: candidates %6.0f/%6.0f  staging %6.0f/%6.0f", m[5] / n, x[5] / n, m[6] / n, x[6] / n);
                fprintf(stderr, "\n");
            }
        }
        cudaFree(d->dbg);
        free(d->dbg_host);
    }
    if(d->par) cudaFree(d->par);
    if(d->rows) cudaFree(d->rows);
    if(d->out) cudaFreeHost(d->out);
    delete d;
    t->track_state = nullptr;
}

int EF_TRACK_FN(device_track_launch)(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    if(!EF_TRACK_FN(device_track_supported)(t))
    {
        t->err = "image too large for the shared-memory candidate store of EF_SOLVE_DEVICE";
        return EF_ERR_UNSUPPORTED;
    }
    BatchArgs B;
    const Layout & lay = d->use_sym ? d->lay_sym : d->lay;
    fill_args(t, d, lay, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3, B.seq[0]);
    void * args[] = {&B};
    const void * fn = d->use_sym ? (d->dbg ? (const void *)k_track_sym<true> : (const void *)k_track_sym<false>)
                                 : (d->dbg ? (const void *)k_track<true> : (const void *)k_track<false>);
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(lay.grid), dim3(kThreads), args, lay.smem_bytes, t->stream);
    t->launches++;
    if(e != cudaSuccess)
    {
        t->err = std::string("cudaLaunchCooperativeKernel: ") + cudaGetErrorString(e);
        return (int)e;
    }
    return EF_OK;
}

int EF_TRACK_FN(device_track_finish)(ef_tracker * t, float * trans, float * rot)
{
    DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
    if(!d) return EF_ERR_BAD_STATE;
    // The solver thread stores the result straight into pinned host memory and raises `status` behind a system-scope
    // fence, so the host takes it from there as soon as it lands instead of waiting for the kernel to retire and the
    // driver to signal the stream (saves the completion-notification latency of every frame).  The stream is still the
    // ordering point for everything that follows on the device; errors and timing modes take the synchronising path.
    cudaError_t e = cudaSuccess;
    const volatile int * flag = &d->out->status;
    if(!t->profile && !d->dbg)
    {
        for(long spins = 0; *flag == 0; ++spins)
        {
            if((spins & 1023) == 1023)
            {
                e = cudaStreamQuery(t->stream);
                if(e != cudaErrorNotReady) break; // finished (the flag is visible by now) or failed
                e = cudaSuccess;
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if(e == cudaSuccess && *flag == 0) e = cudaStreamSynchronize(t->stream);
    }
    else
        e = cudaStreamSynchronize(t->stream);
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
                static const int from[7] = {0, 4, 5, 8, 6, 10, 11}, to[7] = {4, 5, 8, 6, 9, 11, 12};
                for(int ph = 0; ph < 7; ph++)
                {
                    double sum = 0, mx = 0;
                    int cnt = 0;
                    for(int c = 1; c < d->lay.grid; c++)
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

#else // EF_TRACK_GROUPS > 1: the batched build only launches; handles are created, configured and finished by their own variant

// kGroups sequences (handles of one image size, each with its pyramids built and its streams joined) from ONE launch on
// `stream`: group g of every CTA works for ts[g].  The shared memory of a CTA is split evenly, so a group keeps one level's
// lists at a time (no preparation of finer levels ahead).  EF_ERR_UNSUPPORTED when a group's share cannot hold them.
int EF_TRACK_FN(device_track_launch_batch)(ef_tracker * const * ts, const float * const * trans, const float * const * rot, int rgb_only, float icp_weight,
                                           int pyramid, int fast_odom, int so3, cudaStream_t stream)
{
    static bool attr_set[64] = {false}; // per device: function attributes belong to the device's context
    const int dev = ts[0]->device;
    if(dev < 0 || dev >= 64 || !attr_set[dev])
    {
        cudaError_t e = cudaFuncSetAttribute(k_track<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if(e != cudaSuccess) return (int)e;
        if(dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    ef_tracker * t0 = ts[0];
    const int max_grid = t0->num_sms < kMaxGrid ? t0->num_sms : kMaxGrid;
    Layout L;
    if(!compute_layout(t0, max_grid, ((size_t)kMaxDynSmem / kGroups) & ~(size_t)15, L))
    {
        t0->err = "batched launch: image too large for a thread group's share of the shared memory";
        return EF_ERR_UNSUPPORTED;
    }
    BatchArgs B;
    for(int g = 0; g < kGroups; g++)
    {
        ef_tracker * t = ts[g];
        DeviceTrack * d = static_cast<DeviceTrack *>(t->track_state);
        if(!d || t->width != t0->width || t->height != t0->height || t->device != t0->device) return EF_ERR_INVALID_ARGUMENT;
        fill_args(t, d, L, trans[g], rot[g], rgb_only, icp_weight, pyramid, fast_odom, so3, B.seq[g]);
        B.seq[g].dbg = nullptr; // the clock64 trace is a single-launch diagnostic
    }
    void * args[] = {&B};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_track<false>, dim3(L.grid), dim3(kThreads * kGroups), args, L.smem_bytes * kGroups, stream);
    for(int g = 0; g < kGroups; g++) ts[g]->launches++;
    if(e != cudaSuccess)
    {
        t0->err = std::string("cudaLaunchCooperativeKernel (batched): ") + cudaGetErrorString(e);
        return (int)e;
    }
    return EF_OK;
}
#endif

} // namespace ef
