// ef_api.cu -- C ABI (include/ef_track.h): tracker handle (Tier 1) and per-operator wrappers (Tier 2).
//
// Tier 1 restates the control flow of elasticfusionpublic/Core/src/Utils/RGBDOdometry.cpp on top of
// our own kernels.  EF_SOLVE_HOST keeps the reference's structure (one step kernel per operator call,
// 6x6 LDLT + pose update in double on the host) minus its cudaMalloc/cudaFree and double syncs;
// EF_SOLVE_DEVICE hands the whole SO(3) + Gauss-Newton loop to one persistent kernel
// (ef_track_kernel.cu).  No CPU fallback exists: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <new>
#include <utility>

#include "ef_hostmath.h"
#include "ef_kernels.h"
#include "ef_tracker.h"

using namespace ef;

#define EF_API extern "C" __attribute__((visibility("default")))

namespace
{

int fail(ef_tracker * t, int code, const char * what)
{
    if(t)
    {
        char buf[256];
        if(code > 0) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString((cudaError_t)code));
        else snprintf(buf, sizeof(buf), "%s (code %d)", what, code);
        t->err = buf;
    }
    return code;
}

#define EF_CUDA(t, call)                                      \
    do                                                        \
    {                                                         \
        cudaError_t e_ = (call);                              \
        if(e_ != cudaSuccess) return fail((t), (int)e_, #call); \
    } while(0)

#define EF_LAUNCH(t, call)                                    \
    do                                                        \
    {                                                         \
        cudaError_t e_ = (call);                              \
        (t)->launches++;                                      \
        if(e_ != cudaSuccess) return fail((t), (int)e_, #call); \
    } while(0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// A handle lives on the device that was current when it was created (ef_tracker_create); every Tier-1 entry makes that
// device current for the duration of the call and restores the caller's, so one host thread may drive handles on
// several GPUs (the reference is single-device: it never calls cudaSetDevice after start-up).
struct DeviceGuard
{
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const ef_tracker * t)
    {
        if(t && cudaGetDevice(&prev) == cudaSuccess && prev != t->device) switched = cudaSetDevice(t->device) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if(switched) cudaSetDevice(prev);
    }
};
#define EF_ON_DEVICE(t) DeviceGuard ef_device_guard_(t)

int flush_deferred(ef_tracker * t); // EF_OPT_DEFER_BUILD: build what the recorded init* calls asked for

struct ArenaPlan
{
    size_t off = 0;
    size_t take(size_t bytes)
    {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

void level_intr(const ef_tracker * t, int level, float & fx, float & fy, float & cx, float & cy)
{
    const int div = 1 << level; // CameraModel::operator()(level), Cuda/types.cuh:94-98
    fx = t->fx / div;
    fy = t->fy / div;
    cx = t->cx / div;
    cy = t->cy / div;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// handle lifetime
// ------------------------------------------------------------------------------------------------
EF_API int ef_abi_version(void) { return EF_ABI_VERSION; }

EF_API int ef_device_count(void)
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

EF_API float ef_default_dist_thresh(void) { return 0.10f; }
EF_API float ef_default_angle_thresh(void) { return sinf(20.f * 3.14159254f / 180.f); } // RGBDOdometry.h:39

EF_API size_t ef_op_scratch_bytes(void) { return kScratchBytes; }

constexpr int kNumAux = 3;

// pinned result block (floats): one fetched result at 0; graph replay keeps the three results of an iteration and the
// iteration's parameters side by side
constexpr int kHostResultFloats = 256;
constexpr int kHostIterFloats = 128; // fused iteration (ef_iter_fused.cu): 64 sums + sequence word, written by the device (mapped)
constexpr int kHResid = 64, kHRgb = 96, kHIcp = 128, kHIter = 160;
static_assert(kHIter * sizeof(float) + sizeof(IterParams) <= kHostResultFloats * sizeof(float), "pinned result block");

static void destroy_iter_graphs(ef_tracker * t)
{
    for(int p = 0; p < 2; p++)
        for(int i = 0; i < kNumPyrs; i++)
        {
            if(t->iter_graph[p][i]) cudaGraphExecDestroy(t->iter_graph[p][i]);
            t->iter_graph[p][i] = nullptr;
            t->iter_graph_key[p][i] = 0;
        }
}

static void destroy_aux(ef_tracker * t)
{
    for(int i = 0; i < kNumAux; i++)
    {
        if(t->aux[i]) cudaStreamDestroy(t->aux[i]);
        if(t->ev_join[i]) cudaEventDestroy(t->ev_join[i]);
        t->aux[i] = nullptr;
        t->ev_join[i] = nullptr;
    }
    if(t->ev_fork) cudaEventDestroy(t->ev_fork);
    t->ev_fork = nullptr;
    if(t->copy_stream) cudaStreamDestroy(t->copy_stream);
    if(t->ev_stage_free) cudaEventDestroy(t->ev_stage_free);
    if(t->ev_stage_ready) cudaEventDestroy(t->ev_stage_ready);
    t->copy_stream = nullptr;
    t->ev_stage_free = t->ev_stage_ready = nullptr;
}

// Fork/join of independent builders.  The current-frame depth pyramid (aux 0) and the model RGB-D pyramid (aux 1)
// depend on nothing the other builders of a frame produce except what is already enqueued on the handle's stream, so
// they run on internal streams ordered AFTER the handle's stream position at the time of the call (fork) and every
// consumer -- the solve, downloads, builders that write the same buffers -- first waits for them (join).
static cudaStream_t fork_stream(ef_tracker * t, int which)
{
    if(!t->aux_streams) return t->stream;
    if(cudaEventRecord(t->ev_fork, t->stream) != cudaSuccess || cudaStreamWaitEvent(t->aux[which], t->ev_fork, 0) != cudaSuccess) return t->stream;
    return t->aux[which];
}

static void fork_done(ef_tracker * t, int which, cudaStream_t s)
{
    if(s == t->stream) return;
    if(cudaEventRecord(t->ev_join[which], s) == cudaSuccess) t->aux_dirty[which] = true;
    else cudaStreamSynchronize(s);
}

// staging buffers of the _array / _host entry points (see stage_guard)
enum { kStageV = 0, kStageN = 1, kStageModelRgba = 2, kStageDepth = 3, kStageRgba = 4 };

// a builder on internal stream `which` reads `src`: if that is a staging buffer, remember who reads it
static void note_stage_reader(ef_tracker * t, const void * src, int which, cudaStream_t s)
{
    if(s == t->stream) return;
    const void * buf[5] = {t->stage_v, t->stage_n, t->stage_rgba_model, t->stage_depth, t->stage_rgba};
    for(int i = 0; i < 5; i++)
        if(src == buf[i]) t->stage_reader[i] = which;
}

static int join_streams(ef_tracker * t)
{
    for(int i = 0; i < kNumAux; i++)
    {
        if(!t->aux_dirty[i]) continue;
        t->aux_dirty[i] = false;
        EF_CUDA(t, cudaStreamWaitEvent(t->stream, t->ev_join[i], 0));
    }
    return EF_OK;
}

EF_API int ef_tracker_create(int width, int height, float cx, float cy, float fx, float fy, float dist_thresh, float angle_thresh, void * stream,
                             ef_tracker ** out)
{
    // any size the reference accepts (RGBDOdometry.cpp:21-111: level i is (width >> i) x (height >> i)); the coarsest level
    // must still hold the 8 pixels the tracker kernel's clamped window loads assume
    if(!out || width < 16 || height < 16) return EF_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE; // the product has no CPU path

    ef_tracker * t = new(std::nothrow) ef_tracker();
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    t->width = width; t->height = height;
    t->cx = cx; t->cy = cy; t->fx = fx; t->fy = fy;
    t->dist_thresh = dist_thresh; t->angle_thresh = angle_thresh;
    t->sobel_scale = (float)(1.0 / pow(2.0, 3));
    t->max_depth_delta_rgb = 0.07f;
    t->max_depth_rgb = 6.0f;
    t->min_grad[0] = 5; t->min_grad[1] = 3; t->min_grad[2] = 1;
    for(int i = 0; i < kNumPyrs; i++) t->dims[i] = LevelDims{height >> i, width >> i};
    t->solve_mode = EF_SOLVE_DEVICE; // falls back to EF_SOLVE_HOST below when the image does not fit the persistent kernel
    t->use_graph = 0;
    // the fused builders move two / four pixels per load and store: images whose sides are not multiples of 4 take the
    // one-kernel-per-operator builders (any size, any pitch)
    t->vector_ok = (width % 4 == 0) && (height % 4 == 0);
    t->fused_build = t->vector_ok ? 1 : 0;
    for(int i = 0; i < 5; i++) t->stage_reader[i] = -1;
    t->launches = 0;
    t->grid_ctas = 0;
    t->host_fused = 1;
    t->iter_seq = 0;
    t->aux_streams = 1;
    t->frame_build = 2;
    t->defer_build = 0;
    t->deferred.have = 0;
    for(int i = 0; i < kNumAux; i++)
    {
        t->aux[i] = nullptr;
        t->ev_join[i] = nullptr;
        t->aux_dirty[i] = false;
    }
    t->ev_fork = nullptr;
    t->copy_stream = nullptr;
    t->ev_stage_free = t->ev_stage_ready = nullptr;
    t->stage_free_valid = false;
    t->profile = 0;
    t->ev_begin = t->ev_end = nullptr;
    t->ev_pending = false;
    t->prof_ms = 0.0;
    t->prof_calls = 0;
    t->deriv_valid = false;
    t->launch_pending = false;
    t->track_state = nullptr;
    t->h_track_out = nullptr;
    t->graph_scratch = nullptr;
    t->image_parity = 0;
    for(int i = 0; i < kNumPyrs; i++) t->iter_graph[0][i] = t->iter_graph[1][i] = nullptr;
    for(int i = 0; i < kNumPyrs; i++) t->iter_graph_key[0][i] = t->iter_graph_key[1][i] = 0;
    memset(&t->st, 0, sizeof(t->st));
    memset(&t->stages, 0, sizeof(t->stages));
    t->st.last_icp_count = t->st.last_rgb_count = t->st.last_so3_count = (float)(width * height); // RGBDOdometry.cpp:26-31

    cudaError_t e = cudaGetDevice(&t->device);
    if(e == cudaSuccess) e = cudaDeviceGetAttribute(&t->num_sms, cudaDevAttrMultiProcessorCount, t->device);
    if(e != cudaSuccess) { delete t; return (int)e; }

    if(stream) { t->stream = (cudaStream_t)stream; t->own_stream = false; }
    else
    {
        e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
        if(e != cudaSuccess) { delete t; return (int)e; }
        t->own_stream = true;
    }

    for(int i = 0; i < kNumAux && e == cudaSuccess; i++)
    {
        e = cudaStreamCreateWithFlags(&t->aux[i], cudaStreamNonBlocking);
        if(e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_join[i], cudaEventDisableTiming);
    }
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming);
    if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_stage_free, cudaEventDisableTiming);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_stage_ready, cudaEventDisableTiming);
    if(e != cudaSuccess)
    {
        destroy_aux(t);
        if(t->own_stream) cudaStreamDestroy(t->stream);
        delete t;
        return (int)e;
    }

    // arena plan
    ArenaPlan plan;
    size_t o_depth_tmp[kNumPyrs], o_vc[kNumPyrs], o_nc[kNumPyrs], o_vp[kNumPyrs], o_np[kNumPyrs], o_ld[kNumPyrs], o_nd[kNumPyrs], o_li[kNumPyrs],
        o_ni[kNumPyrs], o_lni[kNumPyrs], o_dx[kNumPyrs], o_dy[kNumPyrs], o_cor[kNumPyrs], o_cl[kNumPyrs];
    for(int i = 0; i < kNumPyrs; i++)
    {
        const size_t n = t->dims[i].n();
        o_depth_tmp[i] = plan.take(n * 2);
        o_vc[i] = plan.take(n * 12); o_nc[i] = plan.take(n * 12);
        o_vp[i] = plan.take(n * 12); o_np[i] = plan.take(n * 12);
        o_ld[i] = plan.take(n * 4); o_nd[i] = plan.take(n * 4);
        o_li[i] = plan.take(n); o_ni[i] = plan.take(n); o_lni[i] = plan.take(n);
        o_dx[i] = plan.take(n * 2); o_dy[i] = plan.take(n * 2);
        o_cor[i] = plan.take(n * 16); o_cl[i] = plan.take(n * 12);
    }
    const size_t n0 = t->dims[0].n();
    const size_t o_tmpz = plan.take(n0 * 4);
    const size_t o_filt = plan.take(n0 * 2);
    const size_t o_scratch = plan.take(kScratchBytes);
    const size_t o_scratch2 = plan.take(kScratchBytes);
    const size_t o_sd = plan.take(n0 * 2), o_sr = plan.take(n0 * 4), o_srm = plan.take(n0 * 4), o_sv = plan.take(n0 * 16), o_sn = plan.take(n0 * 16);
    t->arena_bytes = plan.off;

    e = cudaMalloc(&t->arena, t->arena_bytes);
    if(e == cudaSuccess) e = cudaMemsetAsync(t->arena, 0, t->arena_bytes, t->stream);
    if(e == cudaSuccess) e = cudaHostAlloc((void **)&t->h_result, (kHostResultFloats + kHostIterFloats) * sizeof(float), cudaHostAllocMapped);
    if(e == cudaSuccess) memset(t->h_result, 0, (kHostResultFloats + kHostIterFloats) * sizeof(float));
    if(e != cudaSuccess)
    {
        if(t->arena) cudaFree(t->arena);
        destroy_aux(t);
        if(t->own_stream) cudaStreamDestroy(t->stream);
        delete t;
        return (int)e;
    }
    char * base = static_cast<char *>(t->arena);
    for(int i = 0; i < kNumPyrs; i++)
    {
        t->depth_tmp[i] = (uint16_t *)(base + o_depth_tmp[i]);
        t->vmap_curr[i] = (float *)(base + o_vc[i]); t->nmap_curr[i] = (float *)(base + o_nc[i]);
        t->vmap_g_prev[i] = (float *)(base + o_vp[i]); t->nmap_g_prev[i] = (float *)(base + o_np[i]);
        t->last_depth[i] = (float *)(base + o_ld[i]); t->next_depth[i] = (float *)(base + o_nd[i]);
        t->last_image[i] = (uint8_t *)(base + o_li[i]); t->next_image[i] = (uint8_t *)(base + o_ni[i]);
        t->last_next_image[i] = (uint8_t *)(base + o_lni[i]);
        t->dIdx[i] = (int16_t *)(base + o_dx[i]); t->dIdy[i] = (int16_t *)(base + o_dy[i]);
        t->corres[i] = base + o_cor[i];
        t->cloud[i] = (float *)(base + o_cl[i]);
    }
    t->tmp_z = (float *)(base + o_tmpz);
    t->filt_depth = (uint16_t *)(base + o_filt);
    t->scratch = base + o_scratch;
    t->scratch2 = base + o_scratch2;
    t->stage_depth = (uint16_t *)(base + o_sd);
    t->stage_rgba = (uint8_t *)(base + o_sr);
    t->stage_rgba_model = (uint8_t *)(base + o_srm);
    t->stage_v = (float *)(base + o_sv);
    t->stage_n = (float *)(base + o_sn);

    const int rc = device_track_init(t);
    if(rc != EF_OK)
    {
        cudaFree(t->arena);
        cudaFreeHost(t->h_result);
        destroy_aux(t);
        if(t->own_stream) cudaStreamDestroy(t->stream);
        delete t;
        return rc;
    }
    if(!device_track_supported(t)) t->solve_mode = EF_SOLVE_HOST;
    e = cudaStreamSynchronize(t->stream);
    if(e != cudaSuccess) { ef_tracker_destroy(t); return (int)e; }
    *out = t;
    return EF_OK;
}

EF_API int ef_tracker_destroy(ef_tracker * t)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_OK;
    for(int i = 0; i < kNumAux; i++)
        if(t->aux[i]) cudaStreamSynchronize(t->aux[i]);
    cudaStreamSynchronize(t->stream);
    destroy_aux(t);
    if(t->ev_begin) cudaEventDestroy(t->ev_begin);
    if(t->ev_end) cudaEventDestroy(t->ev_end);
    device_track_destroy(t);
    destroy_iter_graphs(t);
    if(t->graph_scratch) cudaFree(t->graph_scratch);
    cudaFree(t->arena);
    cudaFreeHost(t->h_result);
    if(t->own_stream) cudaStreamDestroy(t->stream);
    delete t;
    return EF_OK;
}

EF_API int ef_tracker_set_option(ef_tracker * t, int key, int value)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    switch(key)
    {
    case EF_OPT_SOLVE_MODE:
        if(value != EF_SOLVE_HOST && value != EF_SOLVE_DEVICE) return fail(t, EF_ERR_INVALID_ARGUMENT, "bad solve mode");
        if(t->launch_pending) return fail(t, EF_ERR_BAD_STATE, "a launch is pending");
        if(value == EF_SOLVE_DEVICE && !device_track_supported(t))
            return fail(t, EF_ERR_UNSUPPORTED, "image too large for the shared-memory candidate store of EF_SOLVE_DEVICE");
        t->solve_mode = value;
        return EF_OK;
    case EF_OPT_USE_GRAPH:
        if(value && !t->graph_scratch)
        {
            // three reduction scratch blocks (one per operator of an iteration) + the device copy of the iteration parameters
            const size_t bytes = 3 * kScratchBytes + sizeof(IterParams);
            EF_CUDA(t, cudaMalloc(&t->graph_scratch, bytes));
            EF_CUDA(t, cudaMemsetAsync(t->graph_scratch, 0, bytes, t->stream));
        }
        t->use_graph = value ? 1 : 0;
        return EF_OK;
    case EF_OPT_FUSED_BUILD:
        if(value && !t->vector_ok) return fail(t, EF_ERR_UNSUPPORTED, "fused builders need image sides that are multiples of 4");
        t->fused_build = value ? 1 : 0;
        return EF_OK;
    case EF_OPT_AUX_STREAMS:
    {
        const int rc = join_streams(t);
        if(rc) return rc;
        t->aux_streams = value ? 1 : 0;
        return EF_OK;
    }
    case EF_OPT_FRAME_BUILD:
        if(value < 0 || value > 2) return fail(t, EF_ERR_INVALID_ARGUMENT, "bad frame-build mode");
        t->frame_build = value;
        return EF_OK;
    case EF_OPT_DEFER_BUILD:
    {
        const int rc = flush_deferred(t);
        if(rc) return rc;
        t->defer_build = value ? 1 : 0;
        return EF_OK;
    }
    case EF_OPT_HOST_FUSED: t->host_fused = value ? 1 : 0; return EF_OK;
    case EF_OPT_GRID_CTAS:
        if(t->launch_pending) return fail(t, EF_ERR_BAD_STATE, "a launch is pending");
    {
        const int rc = device_track_configure(t, value);
        if(rc == EF_OK) t->grid_ctas = value;
        return rc;
    }
    case EF_OPT_PROFILE:
        if(value && !t->ev_begin)
        {
            EF_CUDA(t, cudaEventCreate(&t->ev_begin));
            EF_CUDA(t, cudaEventCreate(&t->ev_end));
        }
        t->profile = value ? 1 : 0;
        return EF_OK;
    default: return fail(t, EF_ERR_INVALID_ARGUMENT, "unknown option");
    }
}

EF_API int ef_tracker_get_option(ef_tracker * t, int key, int * value)
{
    if(!t || !value) return EF_ERR_INVALID_ARGUMENT;
    switch(key)
    {
    case EF_OPT_SOLVE_MODE: *value = t->solve_mode; return EF_OK;
    case EF_OPT_USE_GRAPH: *value = t->use_graph; return EF_OK;
    case EF_OPT_FUSED_BUILD: *value = t->fused_build; return EF_OK;
    case EF_OPT_PROFILE: *value = t->profile; return EF_OK;
    case EF_OPT_GRID_CTAS: *value = t->grid_ctas; return EF_OK;
    case EF_OPT_AUX_STREAMS: *value = t->aux_streams; return EF_OK;
    case EF_OPT_FRAME_BUILD: *value = t->frame_build; return EF_OK;
    case EF_OPT_DEFER_BUILD: *value = t->defer_build; return EF_OK;
    case EF_OPT_HOST_FUSED: *value = t->host_fused; return EF_OK;
    default: return EF_ERR_INVALID_ARGUMENT;
    }
}

EF_API const char * ef_last_error(const ef_tracker * t) { return t ? t->err.c_str() : "null handle"; }
EF_API void * ef_tracker_stream(ef_tracker * t) { return t ? (void *)t->stream : nullptr; }
EF_API long long ef_tracker_launch_count(const ef_tracker * t) { return t ? t->launches : 0; }

EF_API int ef_tracker_profile(ef_tracker * t, double * ms, long long * calls)
{
    if(!t || !ms || !calls) return EF_ERR_INVALID_ARGUMENT;
    *ms = t->prof_ms;
    *calls = t->prof_calls;
    t->prof_ms = 0.0;
    t->prof_calls = 0;
    return EF_OK;
}

EF_API int ef_tracker_stage_times(ef_tracker * t, ef_stage_times * out)
{
    if(!t || !out) return EF_ERR_INVALID_ARGUMENT;
    *out = t->stages;
    return EF_OK;
}

EF_API int ef_tracker_trace(ef_tracker * t, double * cycles32, long long * calls)
{
    if(!t || !cycles32 || !calls) return EF_ERR_INVALID_ARGUMENT;
    return device_track_trace(t, cycles32, calls);
}

EF_API int ef_tracker_wait_event(ef_tracker * t, void * cuda_event)
{
    if(!t || !cuda_event) return EF_ERR_INVALID_ARGUMENT;
    EF_ON_DEVICE(t);
    EF_CUDA(t, cudaStreamWaitEvent(t->stream, (cudaEvent_t)cuda_event, 0));
    return EF_OK;
}

EF_API int ef_tracker_wait_stream(ef_tracker * t, void * cuda_stream)
{
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    cudaStream_t producer = (cudaStream_t)cuda_stream;
    if(producer == t->stream) return EF_OK;
    EF_ON_DEVICE(t);
    const cudaError_t q = cudaStreamQuery(producer);
    if(q == cudaSuccess) return EF_OK; // idle: whatever it produced is complete
    if(q != cudaErrorNotReady) return fail(t, (int)q, "cudaStreamQuery(producer stream)");
    EF_CUDA(t, cudaEventRecord(t->ev_fork, producer)); // (ev_fork is re-recorded by every fork: a wait captures the record at the time of the call)
    EF_CUDA(t, cudaStreamWaitEvent(t->stream, t->ev_fork, 0));
    return EF_OK;
}

EF_API int ef_tracker_synchronize(ef_tracker * t)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    {
        const int rc_flush = flush_deferred(t);
        if(rc_flush) return rc_flush;
    }
    const int rc = join_streams(t);
    if(rc) return rc;
    EF_CUDA(t, cudaStreamSynchronize(t->stream));
    return EF_OK;
}

// ------------------------------------------------------------------------------------------------
// pyramid builders
// ------------------------------------------------------------------------------------------------
// RGBDOdometry.cpp:118-142.  raw_max_depth_m > 0: d_depth is the RAW sensor depth and ElasticFusion::filterDepth
// (ElasticFusion.cpp:775-784) runs first, on the same stream, into the handle's own DEPTH_FILTERED buffer.
static int init_icp_depth(ef_tracker * t, const uint16_t * d_depth, size_t pitch_bytes, float depth_cutoff, float raw_max_depth_m)
{
    if(!t || !d_depth) return EF_ERR_INVALID_ARGUMENT;
    if(t->fused_build)
    {
        // one launch per level: vertex map + normal map (+ dense copy of level 0) + bilateral pyrDown to the next level;
        // nothing else a frame builds is read here, so the three launches go to an internal stream (aux 0)
        cudaStream_t s = fork_stream(t, 0);
        note_stage_reader(t, d_depth, 0, s);
        if(raw_max_depth_m > 0.f)
        {
            EF_LAUNCH(t, launch_depth_bilateral(d_depth, pitch_bytes, t->height, t->width, raw_max_depth_m, t->filt_depth, 0, s));
            d_depth = t->filt_depth;
            pitch_bytes = 0;
        }
        for(int i = 0; i < kNumPyrs; ++i)
        {
            float fx, fy, cx, cy;
            level_intr(t, i, fx, fy, cx, cy);
            EF_LAUNCH(t, launch_depth_level(i == 0 ? d_depth : t->depth_tmp[i], i == 0 ? pitch_bytes : 0, t->dims[i].rows, t->dims[i].cols, fx, fy, cx,
                                            cy, depth_cutoff, t->vmap_curr[i], t->nmap_curr[i], i == 0 ? t->depth_tmp[0] : nullptr,
                                            i + 1 < kNumPyrs ? t->depth_tmp[i + 1] : nullptr, s));
        }
        fork_done(t, 0, s);
        return EF_OK;
    }
    cudaStream_t s = t->stream;
    {
        const int rc = join_streams(t);
        if(rc) return rc;
    }
    if(raw_max_depth_m > 0.f)
    {
        EF_LAUNCH(t, launch_depth_bilateral(d_depth, pitch_bytes, t->height, t->width, raw_max_depth_m, t->filt_depth, 0, s));
        d_depth = t->filt_depth;
        pitch_bytes = 0;
    }
    // level 0 is read in place from the caller's buffer (the reference copies the texture into depth_tmp[0])
    const size_t p0 = pitch_bytes ? pitch_bytes : (size_t)t->width * 2;
    const uint16_t * src = d_depth;
    size_t sp = p0;
    // keep a dense copy of level 0 for ef_tracker_download("depth_tmp", 0)
    EF_CUDA(t, cudaMemcpy2DAsync(t->depth_tmp[0], (size_t)t->width * 2, d_depth, p0, (size_t)t->width * 2, t->height, cudaMemcpyDeviceToDevice, s));
    src = t->depth_tmp[0];
    sp = 0;
    for(int i = 1; i < kNumPyrs; ++i)
    {
        EF_LAUNCH(t, launch_pyr_down_u16(src, sp, t->dims[i - 1].rows, t->dims[i - 1].cols, t->depth_tmp[i], 0, s));
        src = t->depth_tmp[i];
        sp = 0;
    }
    for(int i = 0; i < kNumPyrs; ++i)
    {
        float fx, fy, cx, cy;
        level_intr(t, i, fx, fy, cx, cy);
        EF_LAUNCH(t, launch_create_vmap(t->depth_tmp[i], 0, t->dims[i].rows, t->dims[i].cols, fx, fy, cx, cy, depth_cutoff, t->vmap_curr[i], 0, s));
        EF_LAUNCH(t, launch_create_nmap(t->vmap_curr[i], 0, t->dims[i].rows, t->dims[i].cols, t->nmap_curr[i], 0, s));
    }
    return EF_OK;
}

EF_API int ef_init_icp_depth(ef_tracker * t, const uint16_t * d_depth, size_t pitch_bytes, float depth_cutoff)
{
    EF_ON_DEVICE(t);
    if(t && d_depth && t->defer_build && (pitch_bytes == 0 || pitch_bytes == (size_t)t->width * 2))
    {
        t->deferred.depth = d_depth;
        t->deferred.depth_cutoff = depth_cutoff;
        t->deferred.have |= 4u;
        return EF_OK;
    }
    {
        const int rc = flush_deferred(t);
        if(rc) return rc;
    }
    return init_icp_depth(t, d_depth, pitch_bytes, depth_cutoff, 0.f);
}

// ElasticFusion.cpp:309 (filterDepth) + :348 (initICP) in one call
EF_API int ef_init_icp_depth_raw(ef_tracker * t, const uint16_t * d_raw_depth, size_t pitch_bytes, float max_depth_m, float depth_cutoff)
{
    EF_ON_DEVICE(t);
    if(!(max_depth_m > 0.f)) return t ? fail(t, EF_ERR_INVALID_ARGUMENT, "max_depth_m must be positive") : EF_ERR_INVALID_ARGUMENT;
    {
        const int rc_flush = flush_deferred(t);
        if(rc_flush) return rc_flush;
    }
    return init_icp_depth(t, d_raw_depth, pitch_bytes, depth_cutoff, max_depth_m);
}

static int build_maps(ef_tracker * t, const float * d_v, const float * d_n, float ** vmaps, float ** nmaps, const float * R, const float * tv)
{
    cudaStream_t s = t->stream;
    {
        // writes tmp_z (read by the forked model RGB-D builder) and possibly the maps the forked depth builder writes
        const int rc = join_streams(t);
        if(rc) return rc;
    }
    if(t->fused_build)
    {
        EF_LAUNCH(t, launch_build_maps(d_v, d_n, t->height, t->width, vmaps, nmaps, t->tmp_z, R, tv, s));
        return EF_OK;
    }
    // stands for the copy into vmaps_tmp (RGBDOdometry.cpp:150/178): only the z channel is ever read again (:212)
    EF_LAUNCH(t, launch_extract_z(d_v, (int)t->dims[0].n(), t->tmp_z, s));
    EF_LAUNCH(t, launch_copy_maps(d_v, d_n, t->height, t->width, vmaps[0], nmaps[0], 0, s));
    for(int i = 1; i < kNumPyrs; ++i)
    {
        EF_LAUNCH(t, launch_resize_map(vmaps[i - 1], 0, t->dims[i - 1].rows, t->dims[i - 1].cols, vmaps[i], 0, 0, s));
        EF_LAUNCH(t, launch_resize_map(nmaps[i - 1], 0, t->dims[i - 1].rows, t->dims[i - 1].cols, nmaps[i], 0, 1, s));
    }
    if(R)
        for(int i = 0; i < kNumPyrs; ++i)
            EF_LAUNCH(t, launch_transform_maps(vmaps[i], nmaps[i], 0, t->dims[i].rows, t->dims[i].cols, R, tv, vmaps[i], nmaps[i], 0, s));
    return EF_OK;
}

// RGBDOdometry.cpp:144-167
EF_API int ef_init_icp_maps(ef_tracker * t, const float * d_v, const float * d_n, float depth_cutoff)
{
    EF_ON_DEVICE(t);
    (void)depth_cutoff; // unused by the reference as well
    if(!t || !d_v || !d_n) return EF_ERR_INVALID_ARGUMENT;
    {
        const int rc_flush = flush_deferred(t);
        if(rc_flush) return rc_flush;
    }
    return build_maps(t, d_v, d_n, t->vmap_curr, t->nmap_curr, nullptr, nullptr);
}

// RGBDOdometry.cpp:169-206
EF_API int ef_init_icp_model(ef_tracker * t, const float * d_v, const float * d_n, float depth_cutoff, const float * pose)
{
    EF_ON_DEVICE(t);
    (void)depth_cutoff;
    if(!t || !d_v || !d_n || !pose) return EF_ERR_INVALID_ARGUMENT;
    if(t->defer_build)
    {
        t->deferred.v = d_v;
        t->deferred.n = d_n;
        memcpy(t->deferred.pose, pose, sizeof(t->deferred.pose));
        t->deferred.have |= 1u;
        return EF_OK;
    }
    const float R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
    const float tv[3] = {pose[3], pose[7], pose[11]};
    return build_maps(t, d_v, d_n, t->vmap_g_prev, t->nmap_g_prev, R, tv);
}

// RGBDOdometry.cpp:208-235
static int populate_rgbd(ef_tracker * t, const uint8_t * d_rgba, size_t pitch, float ** depths, uint8_t ** images, cudaStream_t s,
                         const float * z = nullptr, int z_stride = 1)
{
    if(t->fused_build)
    {
        if(!z) z = t->tmp_z; // the z channel kept by the last initICP(maps) / initICPModel (stands for vmaps_tmp, RGBDOdometry.cpp:212)
        EF_LAUNCH(t, launch_rgbd_level0(d_rgba, pitch, z, z_stride, t->max_depth_rgb, t->height, t->width, images[0], depths[0], images[1], depths[1], s));
        EF_LAUNCH(t, launch_rgbd_level1(images[1], depths[1], t->dims[1].rows, t->dims[1].cols, images[2], depths[2], s));
        return EF_OK;
    }
    EF_LAUNCH(t, launch_z_to_depth(t->tmp_z, t->height, t->width, t->max_depth_rgb, depths[0], 0, s));
    for(int i = 0; i + 1 < kNumPyrs; i++)
        EF_LAUNCH(t, launch_pyr_down_gauss_f32(depths[i], 0, t->dims[i].rows, t->dims[i].cols, depths[i + 1], 0, s));
    EF_LAUNCH(t, launch_bgr_to_intensity(d_rgba, pitch, t->height, t->width, images[0], 0, s));
    for(int i = 0; i + 1 < kNumPyrs; i++)
        EF_LAUNCH(t, launch_pyr_down_gauss_u8(images[i], 0, t->dims[i].rows, t->dims[i].cols, images[i + 1], 0, s));
    return EF_OK;
}

EF_API int ef_init_rgb(ef_tracker * t, const uint8_t * d_rgba, size_t pitch)
{
    EF_ON_DEVICE(t);
    if(!t || !d_rgba) return EF_ERR_INVALID_ARGUMENT;
    if(t->defer_build && (pitch == 0 || pitch == (size_t)t->width * 4))
    {
        t->deferred.rgba = d_rgba;
        t->deferred.have |= 8u;
        return EF_OK;
    }
    {
        const int rc = flush_deferred(t);
        if(rc) return rc;
    }
    t->deriv_valid = false;
    return populate_rgbd(t, d_rgba, pitch, t->next_depth, t->next_image, t->stream);
}

EF_API int ef_init_rgb_model(ef_tracker * t, const uint8_t * d_rgba, size_t pitch)
{
    EF_ON_DEVICE(t);
    if(!t || !d_rgba) return EF_ERR_INVALID_ARGUMENT;
    if(t->defer_build && (pitch == 0 || pitch == (size_t)t->width * 4))
    {
        t->deferred.model_rgba = d_rgba;
        t->deferred.have |= 2u;
        return EF_OK;
    }
    {
        const int rc = flush_deferred(t);
        if(rc) return rc;
    }
    // reads tmp_z (already enqueued on the handle's stream), writes only the "last" pyramids: internal stream (aux 1)
    cudaStream_t s = t->fused_build ? fork_stream(t, 1) : t->stream;
    note_stage_reader(t, d_rgba, 1, s);
    const int rc = populate_rgbd(t, d_rgba, pitch, t->last_depth, t->last_image, s);
    fork_done(t, 1, s);
    return rc;
}

// RGBDOdometry.cpp:249-265
EF_API int ef_init_first_rgb(ef_tracker * t, const uint8_t * d_rgba, size_t pitch)
{
    EF_ON_DEVICE(t);
    if(!t || !d_rgba) return EF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = t->stream;
    if(t->fused_build)
    {
        EF_LAUNCH(t, launch_rgbd_level0(d_rgba, pitch, nullptr, 1, 0.f, t->height, t->width, t->last_next_image[0], nullptr, t->last_next_image[1], nullptr, s));
        EF_LAUNCH(t, launch_rgbd_level1(t->last_next_image[1], nullptr, t->dims[1].rows, t->dims[1].cols, t->last_next_image[2], nullptr, s));
        return EF_OK;
    }
    EF_LAUNCH(t, launch_bgr_to_intensity(d_rgba, pitch, t->height, t->width, t->last_next_image[0], 0, s));
    for(int i = 0; i + 1 < kNumPyrs; i++)
        EF_LAUNCH(t, launch_pyr_down_gauss_u8(t->last_next_image[i], 0, t->dims[i].rows, t->dims[i].cols, t->last_next_image[i + 1], 0, s));
    return EF_OK;
}

// ---- staging buffers of the _array and _host entry points ----
// Five buffers, one per input of a frame (model vertices, model normals, model colour, depth, colour), so that the calls of
// one frame never share one.  Before a buffer is overwritten
//   (1) a recorded-but-unbuilt init* call (EF_OPT_DEFER_BUILD) that points at it is built, and
//   (2) the handle's stream waits for the internal stream whose builder last read it (the model RGB-D chain and the depth
//       chain run beside the handle's stream, fork_stream),
// so neither a deferred build nor a builder still in flight ever sees the next call's data.
static int stage_guard(ef_tracker * t, int which)
{
    const void * buf[5] = {t->stage_v, t->stage_n, t->stage_rgba_model, t->stage_depth, t->stage_rgba};
    const unsigned have = t->deferred.have;
    const bool referenced = ((have & 1u) && (t->deferred.v == buf[which] || t->deferred.n == buf[which])) ||
                            ((have & 2u) && t->deferred.model_rgba == buf[which]) || ((have & 4u) && t->deferred.depth == buf[which]) ||
                            ((have & 8u) && t->deferred.rgba == buf[which]);
    if(referenced)
    {
        const int rc = flush_deferred(t);
        if(rc) return rc;
    }
    const int r = t->stage_reader[which];
    if(r >= 0 && t->aux_dirty[r]) EF_CUDA(t, cudaStreamWaitEvent(t->stream, t->ev_join[r], 0)); // the join proper still happens later
    t->stage_reader[which] = -1;
    t->stage_free_valid = false; // (the buffer is about to be written and read on the handle's stream: frame_build_all's copy stream orders itself after that)
    return EF_OK;
}

static int stage_array(ef_tracker * t, int which, void * dst, cudaArray_t arr, size_t row_bytes)
{
    if(!arr) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_guard(t, which);
    if(rc) return rc;
    EF_CUDA(t, cudaMemcpy2DFromArrayAsync(dst, row_bytes, arr, 0, 0, row_bytes, t->height, cudaMemcpyDeviceToDevice, t->stream));
    return EF_OK;
}

static int stage_host(ef_tracker * t, int which, void * dst, const void * h, size_t bytes)
{
    const int rc = stage_guard(t, which);
    if(rc) return rc;
    EF_CUDA(t, cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, t->stream));
    return EF_OK;
}

EF_API int ef_init_icp_depth_array(ef_tracker * t, void * arr, float cutoff)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_array(t, kStageDepth, t->stage_depth, (cudaArray_t)arr, (size_t)t->width * 2);
    return rc ? rc : ef_init_icp_depth(t, t->stage_depth, 0, cutoff);
}

EF_API int ef_init_icp_maps_array(ef_tracker * t, void * v, void * n, float cutoff)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    int rc = stage_array(t, kStageV, t->stage_v, (cudaArray_t)v, (size_t)t->width * 16);
    if(!rc) rc = stage_array(t, kStageN, t->stage_n, (cudaArray_t)n, (size_t)t->width * 16);
    return rc ? rc : ef_init_icp_maps(t, t->stage_v, t->stage_n, cutoff);
}

EF_API int ef_init_icp_model_array(ef_tracker * t, void * v, void * n, float cutoff, const float * pose)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    int rc = stage_array(t, kStageV, t->stage_v, (cudaArray_t)v, (size_t)t->width * 16);
    if(!rc) rc = stage_array(t, kStageN, t->stage_n, (cudaArray_t)n, (size_t)t->width * 16);
    return rc ? rc : ef_init_icp_model(t, t->stage_v, t->stage_n, cutoff, pose);
}

EF_API int ef_init_rgb_array(ef_tracker * t, void * arr)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_array(t, kStageRgba, t->stage_rgba, (cudaArray_t)arr, (size_t)t->width * 4);
    return rc ? rc : ef_init_rgb(t, t->stage_rgba, 0);
}

EF_API int ef_init_rgb_model_array(ef_tracker * t, void * arr)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    // the model image has a staging buffer of its own (see ef_init_rgb_model_host)
    const int rc = stage_array(t, kStageModelRgba, t->stage_rgba_model, (cudaArray_t)arr, (size_t)t->width * 4);
    return rc ? rc : ef_init_rgb_model(t, t->stage_rgba_model, 0);
}

EF_API int ef_init_first_rgb_array(ef_tracker * t, void * arr)
{
    EF_ON_DEVICE(t);
    if(!t) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_array(t, kStageRgba, t->stage_rgba, (cudaArray_t)arr, (size_t)t->width * 4);
    return rc ? rc : ef_init_first_rgb(t, t->stage_rgba, 0);
}

// ---- host-buffer variants ----
EF_API int ef_init_icp_depth_host(ef_tracker * t, const uint16_t * h, float cutoff)
{
    EF_ON_DEVICE(t);
    if(!t || !h) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_host(t, kStageDepth, t->stage_depth, h, t->dims[0].n() * 2);
    return rc ? rc : ef_init_icp_depth(t, t->stage_depth, 0, cutoff);
}

EF_API int ef_init_icp_depth_raw_host(ef_tracker * t, const uint16_t * h, float max_depth_m, float cutoff)
{
    EF_ON_DEVICE(t);
    if(!t || !h) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_host(t, kStageDepth, t->stage_depth, h, t->dims[0].n() * 2);
    return rc ? rc : ef_init_icp_depth_raw(t, t->stage_depth, 0, max_depth_m, cutoff);
}

EF_API int ef_init_icp_maps_host(ef_tracker * t, const float * hv, const float * hn, float cutoff)
{
    EF_ON_DEVICE(t);
    if(!t || !hv || !hn) return EF_ERR_INVALID_ARGUMENT;
    int rc = stage_host(t, kStageV, t->stage_v, hv, t->dims[0].n() * 16);
    if(!rc) rc = stage_host(t, kStageN, t->stage_n, hn, t->dims[0].n() * 16);
    return rc ? rc : ef_init_icp_maps(t, t->stage_v, t->stage_n, cutoff);
}

EF_API int ef_init_icp_model_host(ef_tracker * t, const float * hv, const float * hn, float cutoff, const float * pose)
{
    EF_ON_DEVICE(t);
    if(!t || !hv || !hn) return EF_ERR_INVALID_ARGUMENT;
    int rc = stage_host(t, kStageV, t->stage_v, hv, t->dims[0].n() * 16);
    if(!rc) rc = stage_host(t, kStageN, t->stage_n, hn, t->dims[0].n() * 16);
    return rc ? rc : ef_init_icp_model(t, t->stage_v, t->stage_n, cutoff, pose);
}

static int stage_rgba_host(ef_tracker * t, const uint8_t * h) { return stage_host(t, kStageRgba, t->stage_rgba, h, t->dims[0].n() * 4); }

EF_API int ef_init_rgb_host(ef_tracker * t, const uint8_t * h)
{
    EF_ON_DEVICE(t);
    if(!t || !h) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_rgba_host(t, h);
    return rc ? rc : ef_init_rgb(t, t->stage_rgba, 0);
}

EF_API int ef_init_rgb_model_host(ef_tracker * t, const uint8_t * h)
{
    EF_ON_DEVICE(t);
    if(!t || !h) return EF_ERR_INVALID_ARGUMENT;
    // the model image has a staging buffer of its own: its pyramid is built on an internal stream (aux 1) and may still
    // be reading it while the copy of the next image runs on the handle's stream
    const int rc = stage_host(t, kStageModelRgba, t->stage_rgba_model, h, t->dims[0].n() * 4);
    return rc ? rc : ef_init_rgb_model(t, t->stage_rgba_model, 0);
}

EF_API int ef_init_first_rgb_host(ef_tracker * t, const uint8_t * h)
{
    EF_ON_DEVICE(t);
    if(!t || !h) return EF_ERR_INVALID_ARGUMENT;
    const int rc = stage_rgba_host(t, h);
    return rc ? rc : ef_init_first_rgb(t, t->stage_rgba, 0);
}

// ------------------------------------------------------------------------------------------------
// getIncrementalTransformation, host-solve path (RGBDOdometry.cpp:267-603)
// ------------------------------------------------------------------------------------------------
static RgbResArgs rgb_res_args(const ef_tracker * t, int i, const float * krkInv, const float * kt)
{
    RgbResArgs a;
    a.min_scale = (float)(pow(t->min_grad[i], 2.0) / pow(t->sobel_scale, 2.0)); // :442
    a.max_depth_delta = t->max_depth_delta_rgb;
    memcpy(a.kt, kt, sizeof(a.kt)); memcpy(a.krkinv, krkInv, sizeof(a.krkinv));
    a.rows = t->dims[i].rows; a.cols = t->dims[i].cols;
    a.dIdx = t->dIdx[i]; a.dIdy = t->dIdy[i]; a.d_pitch = 0;
    a.last_depth = t->last_depth[i]; a.next_depth = t->next_depth[i]; a.depth_pitch = 0;
    a.last_image = t->last_image[i]; a.next_image = t->next_image[i]; a.image_pitch = 0;
    a.corres = t->corres[i];
    return a;
}

static IcpArgs icp_args(const ef_tracker * t, int i, const float * Rcurr, const float * tcurr, const float * Rprev_inv, const float * tprev)
{
    IcpArgs a;
    memcpy(a.Rcurr, Rcurr, sizeof(a.Rcurr)); memcpy(a.tcurr, tcurr, sizeof(a.tcurr));
    memcpy(a.Rprev_inv, Rprev_inv, sizeof(a.Rprev_inv)); memcpy(a.tprev, tprev, sizeof(a.tprev));
    level_intr(t, i, a.fx, a.fy, a.cx, a.cy);
    a.dist_thresh = t->dist_thresh; a.angle_thresh = t->angle_thresh;
    a.rows = t->dims[i].rows; a.cols = t->dims[i].cols;
    a.vmap_curr = t->vmap_curr[i]; a.nmap_curr = t->nmap_curr[i];
    a.vmap_g_prev = t->vmap_g_prev[i]; a.nmap_g_prev = t->nmap_g_prev[i];
    a.pitch = 0;
    return a;
}

static RgbStepArgs rgb_step_args(const ef_tracker * t, int i, float sigmaVal)
{
    RgbStepArgs a;
    float cx, cy;
    a.corres = t->corres[i];
    a.sigma = sigmaVal;
    a.cloud = t->cloud[i]; a.cloud_pitch = 0;
    level_intr(t, i, a.fx, a.fy, cx, cy);
    a.dIdx = t->dIdx[i]; a.dIdy = t->dIdy[i]; a.d_pitch = 0;
    a.sobel_scale = t->sobel_scale;
    a.rows = t->dims[i].rows; a.cols = t->dims[i].cols;
    return a;
}

// EF_OPT_USE_GRAPH: one Gauss-Newton iteration of level i -- parameter upload, computeRgbResidual, rgbStep (sigma formed on
// the device), icpStep, the three results back to pinned memory -- captured once per handle and level and replayed with ONE
// cudaGraphLaunch and ONE synchronisation per iteration (the plain path: three launches, three copies, three
// synchronisations).  Everything that changes between iterations or frames travels through ef::IterParams; the rest of a
// level (buffers, intrinsics, thresholds) is fixed at handle creation, so the graph is only rebuilt when the set of
// operators (ICP / RGB) changes.
static int iter_graph_for(ef_tracker * t, int i, bool icp, bool rgb)
{
    const int key = 1 + (icp ? 2 : 0) + (rgb ? 4 : 0);
    cudaGraphExec_t & exec = t->iter_graph[t->image_parity][i];
    int & have = t->iter_graph_key[t->image_parity][i];
    if(exec && have == key) return EF_OK;
    if(exec) { cudaGraphExecDestroy(exec); exec = nullptr; have = 0; }
    cudaStream_t s = t->stream;
    char * gs = static_cast<char *>(t->graph_scratch);
    const IterParams * d_it = reinterpret_cast<const IterParams *>(gs + 3 * kScratchBytes);
    const float zero9[9] = {0}, zero3[3] = {0};
    EF_CUDA(t, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaMemcpyAsync(const_cast<IterParams *>(d_it), t->h_result + kHIter, sizeof(IterParams), cudaMemcpyHostToDevice, s);
    // two branches after the parameter upload: the photometric pair (residual -> step, which needs the residual's sums) and the
    // ICP step, which depends on neither -- captured on an internal stream so that the graph runs them side by side
    cudaStream_t si = s;
    if(icp && rgb && t->aux_streams && e == cudaSuccess)
    {
        e = cudaEventRecord(t->ev_fork, s);
        if(e == cudaSuccess) e = cudaStreamWaitEvent(t->aux[0], t->ev_fork, 0);
        si = t->aux[0];
    }
    if(icp)
    {
        if(e == cudaSuccess) e = launch_icp_step(icp_args(t, i, zero9, zero3, zero9, zero3), gs + 2 * kScratchBytes, si, d_it);
        if(e == cudaSuccess)
            e = cudaMemcpyAsync(t->h_result + kHIcp, gs + 2 * kScratchBytes + kScratchResultOff, 29 * sizeof(float), cudaMemcpyDeviceToHost, si);
    }
    if(rgb)
    {
        if(e == cudaSuccess) e = launch_rgb_residual(rgb_res_args(t, i, zero9, zero3), gs, s, d_it);
        if(e == cudaSuccess)
            e = launch_rgb_step(rgb_step_args(t, i, 0.f), gs + kScratchBytes, s, d_it, reinterpret_cast<const int *>(gs + kScratchResultOff));
        if(e == cudaSuccess) e = cudaMemcpyAsync(t->h_result + kHResid, gs + kScratchResultOff, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
        if(e == cudaSuccess)
            e = cudaMemcpyAsync(t->h_result + kHRgb, gs + kScratchBytes + kScratchResultOff, 29 * sizeof(float), cudaMemcpyDeviceToHost, s);
    }
    if(si != s)
    {
        // join (also on failure: an unjoined stream would invalidate the capture's end)
        const cudaError_t ej = cudaEventRecord(t->ev_join[0], si);
        const cudaError_t ew = ej == cudaSuccess ? cudaStreamWaitEvent(s, t->ev_join[0], 0) : ej;
        if(e == cudaSuccess) e = ew;
    }
    cudaGraph_t g = nullptr;
    const cudaError_t e2 = cudaStreamEndCapture(s, &g);
    if(e == cudaSuccess) e = e2;
    if(e == cudaSuccess) e = cudaGraphInstantiate(&exec, g, 0);
    if(g) cudaGraphDestroy(g);
    if(e != cudaSuccess) return fail(t, (int)e, "capturing the iteration graph");
    have = key;
    return EF_OK;
}

// The reference brackets its four step operators with Stopwatch TICK / TOCK (Utils/RGBDOdometry.cpp:333/346 so3Step, :441/458
// computeRgbResidual, :493/512 icpStep, :523/538 rgbStep): wall-clock milliseconds around the blocking call, the LAST pair
// winning (Utils/Stopwatch.h:59-82; GPUTest.cpp:283-286 reads them).  Same keys, same meaning, plus the sums over one
// getIncrementalTransformation.
namespace
{
struct StageTimer
{
    float * last, * sum;
    std::chrono::steady_clock::time_point t0;
    StageTimer(float & l, float & s) : last(&l), sum(&s), t0(std::chrono::steady_clock::now()) {}
    ~StageTimer()
    {
        const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *last = ms;
        *sum += ms;
    }
};
} // namespace

static int fetch_result(ef_tracker * t, int nfloats)
{
    EF_CUDA(t, cudaMemcpyAsync(t->h_result, static_cast<char *>(t->scratch) + kScratchResultOff, nfloats * sizeof(float), cudaMemcpyDeviceToHost,
                               t->stream));
    EF_CUDA(t, cudaStreamSynchronize(t->stream));
    return EF_OK;
}

static int compute_derivatives(ef_tracker * t)
{
    if(t->fused_build)
    {
        int rows[3], cols[3];
        for(int i = 0; i < kNumPyrs; i++) { rows[i] = t->dims[i].rows; cols[i] = t->dims[i].cols; }
        EF_LAUNCH(t, launch_derivatives3(t->next_image, t->dIdx, t->dIdy, rows, cols, t->stream));
        t->deriv_valid = true;
        return EF_OK;
    }
    for(int i = 0; i < kNumPyrs; i++) // RGBDOdometry.cpp:284-290
        EF_LAUNCH(t, launch_derivative_images(t->next_image[i], 0, t->dims[i].rows, t->dims[i].cols, t->dIdx[i], t->dIdy[i], 0, t->stream));
    t->deriv_valid = true;
    return EF_OK;
}

static int track_host(ef_tracker * t, float * trans, float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3)
{
    cudaStream_t s = t->stream;
    {
        const int rc = join_streams(t);
        if(rc) return rc;
    }
    const bool icp = !rgb_only && icp_weight > 0; // :275
    const bool rgb = rgb_only || icp_weight < 100; // :276

    float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
    memcpy(Rprev, rot, sizeof(Rprev)); memcpy(tprev, trans, sizeof(tprev));
    memcpy(Rcurr, rot, sizeof(Rcurr)); memcpy(tcurr, trans, sizeof(tcurr));

    if(rgb)
    {
        const int rc = compute_derivatives(t);
        if(rc) return rc;
    }
    // EF_OPT_HOST_FUSED: two launches per iteration, results through mapped pinned memory (ef_iter_fused.cu).  The 8-byte
    // correspondences and the gate image of a level live where the plain path keeps its 16-byte DataTerm records.
    const bool fused = t->host_fused && !t->use_graph && t->width % 16 == 0 && t->width < 4096 && t->height < 4096;
    auto rec0_of = [&](int i) { return reinterpret_cast<unsigned *>(t->corres[i]); };
    auto rec1_of = [&](int i) { return reinterpret_cast<float *>(t->corres[i]) + t->dims[i].n(); };
    auto gate_of = [&](int i) { return reinterpret_cast<float *>(t->corres[i]) + 2 * t->dims[i].n(); };
    float * const h_iter = t->h_result + kHostResultFloats;
    if(fused && rgb)
    {
        HmGatesArgs g;
        for(int i = 0; i < kNumPyrs; i++)
        {
            g.dIdx[i] = t->dIdx[i]; g.dIdy[i] = t->dIdy[i];
            g.next_depth[i] = t->next_depth[i]; g.next_image[i] = t->next_image[i];
            g.gate_depth[i] = gate_of(i);
            g.min_scale[i] = (float)(pow(t->min_grad[i], 2.0) / pow(t->sobel_scale, 2.0)); // :442
            g.rows[i] = t->dims[i].rows; g.cols[i] = t->dims[i].cols;
        }
        EF_LAUNCH(t, launch_hm_gates(g, s));
    }

    double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    t->st.so3_iterations = 0;
    t->st.se3_iterations[0] = t->st.se3_iterations[1] = t->st.se3_iterations[2] = 0;

    if(so3) // :294-382
    {
        const int lvl = 2;
        float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        float fx, fy, cx, cy;
        level_intr(t, lvl, fx, fy, cx, cy);
        const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
        double K_inv[9];
        hm::inverse33(K, K_inv);
        float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
        double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

        for(int i = 0; i < 10; i++)
        {
            double KR[9], H[9];
            hm::mul33(K, resultR, KR);
            hm::mul33(KR, K_inv, H);
            So3Args a;
            a.last_image = t->last_next_image[lvl];
            a.next_image = t->next_image[lvl];
            a.image_pitch = 0;
            for(int k = 0; k < 9; k++) { a.image_basis[k] = (float)H[k]; a.kinv[k] = (float)K_inv[k]; a.krlr[k] = (float)KR[k]; }
            a.rows = t->dims[lvl].rows; a.cols = t->dims[lvl].cols;
            int rc;
            {
                StageTimer tick(t->stages.so3_step_ms, t->stages.so3_step_sum_ms);
                EF_LAUNCH(t, launch_so3_step(a, t->scratch, s));
                rc = fetch_result(t, 11);
            }
            if(rc) return rc;
            float jtj[9], jtr[3], residual[2];
            hm::unpack_so3(t->h_result, jtj, jtr, residual);
            t->st.so3_iterations++;

            t->st.last_so3_error = sqrtf(residual[0]) / residual[1]; // :348
            t->st.last_so3_count = residual[1];
            if(t->st.last_so3_error < lastError && lastCount == t->st.last_so3_count) break; // :352
            else if(t->st.last_so3_error > lastError + 0.001)                                 // :356
            {
                t->st.last_so3_error = lastError;
                t->st.last_so3_count = lastCount;
                memcpy(resultR, lastResultR, sizeof(resultR));
                break;
            }
            lastError = t->st.last_so3_error;
            lastCount = t->st.last_so3_count;
            memcpy(lastResultR, resultR, sizeof(resultR));

            float delta[3];
            hm::ldlt_solve<float, 3>(jtj, jtr, delta); // :368
            const double dd[3] = {delta[0], delta[1], delta[2]};
            double rotUpdate[9];
            hm::rodrigues(dd, rotUpdate);
            float ru[9];
            for(int k = 0; k < 9; k++) ru[k] = (float)rotUpdate[k];
            hm::mul33(ru, R_lr, R_lr); // :372
            for(int k = 0; k < 9; k++) resultR[k] = R_lr[k];
        }
    }

    const int iterations[kNumPyrs] = {fast_odom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0}; // :384-386

    float Rprev_inv[9];
    hm::inverse33(Rprev, Rprev_inv); // :388

    double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if(so3)
        for(int x = 0; x < 3; x++)
            for(int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

    for(int i = kNumPyrs - 1; i >= 0; i--) // :405
    {
        const int rows = t->dims[i].rows, cols = t->dims[i].cols;
        float fx, fy, cx, cy;
        level_intr(t, i, fx, fy, cx, cy);
        if(rgb && iterations[i] > 0 && !fused) // :409 (skipped when the level runs no iteration: its only consumer is rgbStep)
            EF_LAUNCH(t, launch_project_points(t->last_depth[i], 0, rows, cols, fx, fy, cx, cy, t->cloud[i], 0, s));

        const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
        double K_inv[9];
        hm::inverse33(K, K_inv);
        t->st.last_rgb_error = FLT_MAX; // :420

        for(int j = 0; j < iterations[i]; j++)
        {
            float krkInv[9], kt[3];
            hm::rgb_warp_params(resultRt, K, K_inv, krkInv, kt); // :424-434

            int sigma = 0, rgbSize = 0;
            const float * icp_sums = t->h_result, * rgb_sums = t->h_result;
            if(t->use_graph)
            {
                // one replayed graph = the whole evaluation of this iteration
                int rc = iter_graph_for(t, i, icp, rgb);
                if(rc) return rc;
                IterParams * h = reinterpret_cast<IterParams *>(t->h_result + kHIter);
                memcpy(h->Rcurr, Rcurr, sizeof(Rcurr)); memcpy(h->tcurr, tcurr, sizeof(tcurr));
                memcpy(h->Rprev_inv, Rprev_inv, sizeof(Rprev_inv)); memcpy(h->tprev, tprev, sizeof(tprev));
                memcpy(h->krkinv, krkInv, sizeof(krkInv)); memcpy(h->kt, kt, sizeof(kt));
                h->rgb_only = rgb_only ? 1 : 0;
                {
                    // (one replayed graph = all operators of the iteration: its time goes to the iteration key)
                    StageTimer tick(t->stages.iteration_ms, t->stages.iteration_sum_ms);
                    EF_CUDA(t, cudaGraphLaunch(t->iter_graph[t->image_parity][i], s));
                    t->launches += (rgb ? 2 : 0) + (icp ? 1 : 0);
                    EF_CUDA(t, cudaStreamSynchronize(s));
                }
                if(rgb)
                {
                    rgbSize = reinterpret_cast<const int *>(t->h_result + kHResid)[0];
                    sigma = reinterpret_cast<const int *>(t->h_result + kHResid)[1];
                }
                icp_sums = t->h_result + kHIcp;
                rgb_sums = t->h_result + kHRgb;
            }
            else if(fused)
            {
                StageTimer tick(t->stages.iteration_ms, t->stages.iteration_sum_ms);
                if(rgb) EF_LAUNCH(t, launch_hm_residual(rgb_res_args(t, i, krkInv, kt), gate_of(i), rec0_of(i), rec1_of(i), t->scratch, s));
                HmStepArgs a;
                a.icp = icp_args(t, i, Rcurr, tcurr, Rprev_inv, tprev);
                a.sobel_scale = t->sobel_scale;
                a.rec0 = rec0_of(i); a.rec1 = rec1_of(i);
                a.next_image = t->next_image[i];
                a.dIdx = t->dIdx[i]; a.dIdy = t->dIdy[i];
                a.residual = reinterpret_cast<const int *>(static_cast<char *>(t->scratch) + kScratchResultOff);
                a.do_icp = icp ? 1 : 0; a.do_rgb = rgb ? 1 : 0; a.rgb_only = rgb_only ? 1 : 0;
                a.h_out = h_iter;
                a.seq = ++t->iter_seq;
                EF_LAUNCH(t, launch_hm_step(a, t->scratch2, s));
                // the last block stores the sums and then the sequence number behind a system-scope fence: poll it
                const volatile unsigned * flag = reinterpret_cast<const volatile unsigned *>(h_iter + 64);
                cudaError_t e = cudaSuccess;
                for(long spins = 0; *flag != a.seq; ++spins)
                {
                    if((spins & 4095) == 4095)
                    {
                        e = cudaStreamQuery(s);
                        if(e != cudaErrorNotReady) break;
                        e = cudaSuccess;
                    }
#if defined(__x86_64__) || defined(__i386__)
                    __builtin_ia32_pause();
#endif
                }
                __atomic_thread_fence(__ATOMIC_ACQUIRE);
                if(e == cudaSuccess && *flag != a.seq) e = cudaStreamSynchronize(s);
                if(e != cudaSuccess || *flag != a.seq) return fail(t, e != cudaSuccess ? (int)e : EF_ERR_BAD_STATE, "fused iteration produced no result");
                rgbSize = reinterpret_cast<const int *>(h_iter)[62];
                sigma = reinterpret_cast<const int *>(h_iter)[63];
                icp_sums = h_iter;
                rgb_sums = h_iter + 32;
            }
            else if(rgb)
            {
                StageTimer tick(t->stages.rgb_residual_ms, t->stages.rgb_residual_sum_ms);
                EF_LAUNCH(t, launch_rgb_residual(rgb_res_args(t, i, krkInv, kt), t->scratch, s));
                const int rc = fetch_result(t, 2);
                if(rc) return rc;
                rgbSize = reinterpret_cast<int *>(t->h_result)[0];
                sigma = reinterpret_cast<int *>(t->h_result)[1];
            }

            float sigmaVal = (float)sqrt((double)(((float)sigma / (float)rgbSize == 0) ? 1 : rgbSize)); // :461 (precedence quirk kept)
            const float rgbError = (float)(sqrt((double)sigma) / (rgbSize == 0 ? 1 : rgbSize));        // :462

            if(rgb_only && rgbError > t->st.last_rgb_error) break; // :464
            t->st.last_rgb_error = rgbError;
            t->st.last_rgb_count = (float)rgbSize;
            if(rgb_only) sigmaVal = -1; // :472

            double A_icp[36] = {0}, b_icp[6] = {0}, A_rgb[36] = {0}, b_rgb[6] = {0};
            if(icp)
            {
                if(!t->use_graph && !fused)
                {
                    StageTimer tick(t->stages.icp_step_ms, t->stages.icp_step_sum_ms);
                    EF_LAUNCH(t, launch_icp_step(icp_args(t, i, Rcurr, tcurr, Rprev_inv, tprev), t->scratch, s));
                    const int rc = fetch_result(t, 29);
                    if(rc) return rc;
                }
                float residual[2];
                hm::unpack_se3(icp_sums, A_icp, b_icp, residual);
                // :515-516.  (When !icp the reference reads `residual` uninitialised; we keep the previous values.)
                t->st.last_icp_error = sqrtf(residual[0]) / residual[1];
                t->st.last_icp_count = residual[1];
            }
            if(rgb)
            {
                if(!t->use_graph && !fused)
                {
                    StageTimer tick(t->stages.rgb_step_ms, t->stages.rgb_step_sum_ms);
                    EF_LAUNCH(t, launch_rgb_step(rgb_step_args(t, i, sigmaVal), t->scratch, s));
                    const int rc = fetch_result(t, 29);
                    if(rc) return rc;
                }
                hm::unpack_se3(rgb_sums, A_rgb, b_rgb, (float *)nullptr);
            }

            double * lastA = t->st.last_A, * lastb = t->st.last_b, result[6];
            if(icp && rgb) // :547-553
            {
                const double w = icp_weight;
                for(int k = 0; k < 36; k++) lastA[k] = A_rgb[k] + w * w * A_icp[k];
                for(int k = 0; k < 6; k++) lastb[k] = b_rgb[k] + w * b_icp[k];
            }
            else if(icp)
            {
                memcpy(lastA, A_icp, sizeof(A_icp));
                memcpy(lastb, b_icp, sizeof(b_icp));
            }
            else
            {
                memcpy(lastA, A_rgb, sizeof(A_rgb));
                memcpy(lastb, b_rgb, sizeof(b_rgb));
            }
            hm::ldlt_solve<double, 6>(lastA, lastb, result);
            t->st.se3_iterations[i]++;

            hm::update_se3(resultRt, result);                      // :573
            hm::compose_pose(resultRt, Rprev, tprev, Rcurr, tcurr); // :575-583
        }
    }

    if(rgb) // :587-591
    {
        const float d[3] = {tcurr[0] - tprev[0], tcurr[1] - tprev[1], tcurr[2] - tprev[2]};
        if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
        {
            memcpy(Rcurr, Rprev, sizeof(Rcurr));
            memcpy(tcurr, tprev, sizeof(tcurr));
        }
    }

    memcpy(trans, tcurr, sizeof(tcurr));
    memcpy(rot, Rcurr, sizeof(Rcurr));
    return EF_OK;
}

static void swap_so3_images(ef_tracker * t)
{
    for(int i = 0; i < kNumPyrs; i++) std::swap(t->last_next_image[i], t->next_image[i]); // RGBDOdometry.cpp:593-599
    t->image_parity ^= 1;
    t->deriv_valid = false;
}

EF_API int ef_get_incremental_transformation_launch(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight,
                                                    int pyramid, int fast_odom, int so3)
{
    EF_ON_DEVICE(t);
    if(!t || !trans || !rot) return EF_ERR_INVALID_ARGUMENT;
    if(t->launch_pending) return fail(t, EF_ERR_BAD_STATE, "a launch is already pending");
    {
        const int rc = flush_deferred(t);
        if(rc) return rc;
    }
    const bool icp = !rgb_only && icp_weight > 0, rgb = rgb_only || icp_weight < 100;
    if(!icp && !rgb) return fail(t, EF_ERR_INVALID_ARGUMENT, "neither ICP nor RGB selected"); // the reference asserts (:566-569)
    memcpy(t->pending.trans, trans, sizeof(t->pending.trans));
    memcpy(t->pending.rot, rot, sizeof(t->pending.rot));
    t->pending.rgb_only = rgb_only; t->pending.icp_weight = icp_weight; t->pending.pyramid = pyramid;
    t->pending.fast_odom = fast_odom; t->pending.so3 = so3;
    t->stages.so3_step_sum_ms = t->stages.rgb_residual_sum_ms = t->stages.icp_step_sum_ms = t->stages.rgb_step_sum_ms = 0.f;
    t->stages.iteration_sum_ms = 0.f;
    t->stages.solve_mode = t->solve_mode;
    t->call_begin = std::chrono::steady_clock::now();
    if(t->solve_mode == EF_SOLVE_DEVICE)
    {
        // computeDerivativeImages (RGBDOdometry.cpp:436-440) is fused into the tracker kernel: every worker CTA derives
        // dIdx / dIdy of its own pixels while it builds its candidate list (device_track_launch reads !deriv_valid).
        // Levels without iterations (pyramid = 0) keep the stand-alone kernel so that all three levels are valid afterwards.
        if(rgb && !t->deriv_valid && !pyramid)
        {
            const int rc = compute_derivatives(t);
            if(rc) return rc;
        }
        {
            const int rc = join_streams(t);
            if(rc) return rc;
        }
        if(t->profile) EF_CUDA(t, cudaEventRecord(t->ev_begin, t->stream));
        const int rc = device_track_launch(t, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
        if(rc) return rc;
        if(rgb) t->deriv_valid = true;
        if(t->profile)
        {
            EF_CUDA(t, cudaEventRecord(t->ev_end, t->stream));
            t->ev_pending = true;
        }
    }
    t->launch_pending = true;
    return EF_OK;
}

EF_API int ef_get_incremental_transformation_finish(ef_tracker * t, float * trans, float * rot, ef_track_stats * stats)
{
    EF_ON_DEVICE(t);
    if(!t || !trans || !rot) return EF_ERR_INVALID_ARGUMENT;
    if(!t->launch_pending) return fail(t, EF_ERR_BAD_STATE, "no launch pending");
    t->launch_pending = false;
    int rc;
    if(t->solve_mode == EF_SOLVE_DEVICE) rc = device_track_finish(t, trans, rot);
    else
    {
        memcpy(trans, t->pending.trans, sizeof(t->pending.trans));
        memcpy(rot, t->pending.rot, sizeof(t->pending.rot));
        if(t->profile) EF_CUDA(t, cudaEventRecord(t->ev_begin, t->stream));
        rc = track_host(t, trans, rot, t->pending.rgb_only, t->pending.icp_weight, t->pending.pyramid, t->pending.fast_odom, t->pending.so3);
        if(!rc && t->profile)
        {
            EF_CUDA(t, cudaEventRecord(t->ev_end, t->stream));
            EF_CUDA(t, cudaEventSynchronize(t->ev_end));
            t->ev_pending = true;
        }
    }
    if(rc) return rc;
    t->stages.call_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t->call_begin).count();
    if(t->ev_pending)
    {
        float ms = 0.f;
        EF_CUDA(t, cudaEventElapsedTime(&ms, t->ev_begin, t->ev_end));
        t->prof_ms += ms;
        t->prof_calls++;
        t->ev_pending = false;
    }
    if(t->pending.so3) swap_so3_images(t);
    if(stats) *stats = t->st;
    return EF_OK;
}

EF_API int ef_get_incremental_transformation(ef_tracker * t, float * trans, float * rot, int rgb_only, float icp_weight, int pyramid,
                                             int fast_odom, int so3, ef_track_stats * stats)
{
    EF_ON_DEVICE(t);
    const int rc = ef_get_incremental_transformation_launch(t, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
    if(rc) return rc;
    return ef_get_incremental_transformation_finish(t, trans, rot, stats);
}

// every pyramid of a frame-to-model frame from one launch (k_build_frame); device inputs, dense rows, 16-byte aligned
static int build_frame(ef_tracker * t, const float * v4, const float * n4, const uint8_t * model_rgba, const uint16_t * depth, const uint8_t * rgba,
                       const float * pose, float depth_cutoff)
{
    ef::FrameBuildArgs a;
    a.rows = t->height;
    a.cols = t->width;
    a.v4 = v4;
    a.n4 = n4;
    a.model_rgba = model_rgba;
    a.depth = depth;
    a.rgba = rgba;
    a.depth_pitch_bytes = 0;
    const float R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
    memcpy(a.R, R, sizeof(R));
    a.t[0] = pose[3]; a.t[1] = pose[7]; a.t[2] = pose[11];
    a.depth_cutoff = depth_cutoff;
    a.rgb_depth_cutoff = t->max_depth_rgb;
    a.tmp_z = t->tmp_z;
    for(int i = 0; i < kNumPyrs; ++i)
    {
        level_intr(t, i, a.fx[i], a.fy[i], a.cx[i], a.cy[i]);
        a.vmap_g_prev[i] = t->vmap_g_prev[i];
        a.nmap_g_prev[i] = t->nmap_g_prev[i];
        a.vmap_curr[i] = t->vmap_curr[i];
        a.nmap_curr[i] = t->nmap_curr[i];
        a.depth_pyr[i] = t->depth_tmp[i];
        a.next_image[i] = t->next_image[i];
        a.last_image[i] = t->last_image[i];
        a.next_depth[i] = t->next_depth[i];
        a.last_depth[i] = t->last_depth[i];
    }
    t->deriv_valid = false;
    EF_LAUNCH(t, launch_build_frame(a, t->stream));
    return EF_OK;
}

// EF_OPT_DEFER_BUILD: the init* calls only recorded their arguments; build now.  All four present (the frameToModel
// sequence, ElasticFusion.cpp:343-368): one launch.  Otherwise the recorded calls run through their own builders in the
// reference's order (initICP* before initRGB*).
namespace
{
int flush_deferred(ef_tracker * t)
{
    if(!t || !t->deferred.have) return EF_OK;
    const unsigned have = t->deferred.have;
    t->deferred.have = 0;
    const int defer = t->defer_build;
    t->defer_build = 0; // the calls below must build
    int rc = EF_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(t->deferred.v) | reinterpret_cast<uintptr_t>(t->deferred.n) |
                           reinterpret_cast<uintptr_t>(t->deferred.model_rgba) | reinterpret_cast<uintptr_t>(t->deferred.depth) |
                           reinterpret_cast<uintptr_t>(t->deferred.rgba)) & 15) == 0;
    if(have == 15u && t->fused_build && aligned)
    {
        rc = join_streams(t);
        if(!rc) rc = build_frame(t, t->deferred.v, t->deferred.n, t->deferred.model_rgba, t->deferred.depth, t->deferred.rgba, t->deferred.pose,
                                 t->deferred.depth_cutoff);
    }
    else
    {
        if(!rc && (have & 1u)) rc = ef_init_icp_model(t, t->deferred.v, t->deferred.n, t->deferred.depth_cutoff, t->deferred.pose);
        if(!rc && (have & 2u)) rc = ef_init_rgb_model(t, t->deferred.model_rgba, 0);
        if(!rc && (have & 4u)) rc = ef_init_icp_depth(t, t->deferred.depth, 0, t->deferred.depth_cutoff);
        if(!rc && (have & 8u)) rc = ef_init_rgb(t, t->deferred.rgba, 0);
    }
    t->defer_build = defer;
    return rc;
}
} // namespace

// ElasticFusion.cpp:343-368 in one call
// Chain on an internal stream that starts at the fork point recorded in t->ev_fork (no new record: all chains of
// one ef_track_frame_to_model call hang off the same point of the handle's stream)
static cudaStream_t fork_from_recorded(ef_tracker * t, int which)
{
    if(cudaStreamWaitEvent(t->aux[which], t->ev_fork, 0) != cudaSuccess) return t->stream;
    return t->aux[which];
}

// the four builders of a frameToModel frame (initICPModel, initRGBModel, initICP, initRGB) for inputs given all at once
static int frame_build_all(ef_tracker * t, const ef_frame_inputs * in, const float * pose)
{
    if(!t || !in || !pose) return EF_ERR_INVALID_ARGUMENT;
    if(!in->vertices_rgba32f || !in->normals_rgba32f || !in->model_rgba8 || !in->depth || !in->rgba8) return EF_ERR_INVALID_ARGUMENT;
    {
        const int rc_flush = flush_deferred(t);
        if(rc_flush) return rc_flush;
    }
    int rc = EF_OK;
    if(in->on_host < 0 || in->on_host > 2) return fail(t, EF_ERR_INVALID_ARGUMENT, "ef_frame_inputs.on_host must be 0, 1 or 2");
    const bool host_model = in->on_host == 1, host_sensor = in->on_host >= 1;
    // (k_build_frame reads its inputs with 8- and 16-byte loads: device inputs that are not 16-byte aligned take the chained builders)
    uintptr_t dev_bits = 0;
    if(!host_model)
        dev_bits |= reinterpret_cast<uintptr_t>(in->vertices_rgba32f) | reinterpret_cast<uintptr_t>(in->normals_rgba32f) |
                    reinterpret_cast<uintptr_t>(in->model_rgba8);
    if(!host_sensor) dev_bits |= reinterpret_cast<uintptr_t>(in->depth) | reinterpret_cast<uintptr_t>(in->rgba8);
    const bool aligned = (dev_bits & 15) == 0;
    if(t->fused_build && aligned && (host_sensor ? t->frame_build == 2 : t->frame_build >= 1))
    {
        // k_build_frame (ef_build_fused.cu): all five inputs are known at once, so every pyramid of the frame comes from
        // one launch -- no chained kernels, no stream forks and joins; the tracker kernel follows on the same stream.
        rc = join_streams(t);
        if(rc) return rc;
        const float * v4 = static_cast<const float *>(in->vertices_rgba32f);
        const float * n4 = static_cast<const float *>(in->normals_rgba32f);
        const uint8_t * mrgba = static_cast<const uint8_t *>(in->model_rgba8);
        const uint16_t * depth = static_cast<const uint16_t *>(in->depth);
        const uint8_t * rgba = static_cast<const uint8_t *>(in->rgba8);
        const size_t n = t->dims[0].n();
        // (join_streams above: no builder of an earlier call still reads a staging buffer)
        if(host_model)
        {
            EF_CUDA(t, cudaMemcpyAsync(t->stage_v, v4, n * 16, cudaMemcpyHostToDevice, t->stream));
            EF_CUDA(t, cudaMemcpyAsync(t->stage_n, n4, n * 16, cudaMemcpyHostToDevice, t->stream));
            EF_CUDA(t, cudaMemcpyAsync(t->stage_rgba_model, mrgba, n * 4, cudaMemcpyHostToDevice, t->stream));
            v4 = t->stage_v; n4 = t->stage_n; mrgba = t->stage_rgba_model;
        }
        if(host_sensor && !host_model && t->copy_stream)
        {
            // The production data flow: the model maps are already on the device -- typically the output of a prediction that is
            // still running on this handle's stream -- and only the sensor frame comes from the host.  Its two copies do not
            // depend on anything enqueued here except the last builder that READ the staging buffers, so they run on their own
            // stream beside the prediction; the builder waits for them.
            if(t->stage_free_valid) EF_CUDA(t, cudaStreamWaitEvent(t->copy_stream, t->ev_stage_free, 0));
            else
            {
                EF_CUDA(t, cudaEventRecord(t->ev_stage_free, t->stream)); // unknown history: order after everything enqueued so far
                EF_CUDA(t, cudaStreamWaitEvent(t->copy_stream, t->ev_stage_free, 0));
            }
            EF_CUDA(t, cudaMemcpyAsync(t->stage_depth, depth, n * 2, cudaMemcpyHostToDevice, t->copy_stream));
            EF_CUDA(t, cudaMemcpyAsync(t->stage_rgba, rgba, n * 4, cudaMemcpyHostToDevice, t->copy_stream));
            EF_CUDA(t, cudaEventRecord(t->ev_stage_ready, t->copy_stream));
            EF_CUDA(t, cudaStreamWaitEvent(t->stream, t->ev_stage_ready, 0));
            depth = t->stage_depth; rgba = t->stage_rgba;
        }
        else if(host_sensor)
        {
            EF_CUDA(t, cudaMemcpyAsync(t->stage_depth, depth, n * 2, cudaMemcpyHostToDevice, t->stream));
            EF_CUDA(t, cudaMemcpyAsync(t->stage_rgba, rgba, n * 4, cudaMemcpyHostToDevice, t->stream));
            depth = t->stage_depth; rgba = t->stage_rgba;
            t->stage_free_valid = false;
        }
        rc = build_frame(t, v4, n4, mrgba, depth, rgba, pose, in->depth_cutoff);
        if(rc) return rc;
        if(host_sensor && !host_model && t->copy_stream)
        {
            EF_CUDA(t, cudaEventRecord(t->ev_stage_free, t->stream)); // the builder that read the two staging buffers
            t->stage_free_valid = true;
        }
    }
    else if(t->fused_build && t->aux_streams && in->on_host == 0)
    {
        // (Host inputs keep the five-call order below: there the copies are the bound, and builders that trickle in between
        // them delay the cooperative launches of the OTHER handles taking turns on the GPU -- measured 3 640 -> 3 040 frames/s.)
        // All five inputs are known at once, so nothing has to wait for anything: the depth of the model (needed by both
        // RGB-D pyramids, RGBDOdometry.cpp:212) is read from the vertex texture itself instead of the copy initICPModel
        // keeps, and the four builder chains run side by side, the longest ones enqueued first:
        //   handle stream  current RGB-D pyramid (2 kernels)      aux 0  current depth -> vertex/normal pyramid (3 kernels)
        //   aux 1          model RGB-D pyramid (2 kernels)        aux 2  model vertex/normal pyramids (1 kernel)
        const float * v4 = static_cast<const float *>(in->vertices_rgba32f);
        const float * n4 = static_cast<const float *>(in->normals_rgba32f);
        const uint8_t * mrgba = static_cast<const uint8_t *>(in->model_rgba8);
        const uint16_t * depth = static_cast<const uint16_t *>(in->depth);
        const uint8_t * rgba = static_cast<const uint8_t *>(in->rgba8);
        rc = join_streams(t);
        if(rc) return rc;
        {
            EF_CUDA(t, cudaEventRecord(t->ev_fork, t->stream));
            // handle stream: current RGB-D pyramid
            t->deriv_valid = false;
            rc = populate_rgbd(t, rgba, 0, t->next_depth, t->next_image, t->stream, v4 + 2, 4);
            if(rc) return rc;
            {
                cudaStream_t s0 = fork_from_recorded(t, 0);
                for(int i = 0; i < kNumPyrs; ++i)
                {
                    float fx, fy, cx, cy;
                    level_intr(t, i, fx, fy, cx, cy);
                    EF_LAUNCH(t, launch_depth_level(i == 0 ? depth : t->depth_tmp[i], 0, t->dims[i].rows, t->dims[i].cols, fx, fy, cx, cy,
                                                    in->depth_cutoff, t->vmap_curr[i], t->nmap_curr[i], i == 0 ? t->depth_tmp[0] : nullptr,
                                                    i + 1 < kNumPyrs ? t->depth_tmp[i + 1] : nullptr, s0));
                }
                fork_done(t, 0, s0);
            }
            {
                cudaStream_t s1 = fork_from_recorded(t, 1);
                rc = populate_rgbd(t, mrgba, 0, t->last_depth, t->last_image, s1, v4 + 2, 4);
                fork_done(t, 1, s1);
                if(rc) return rc;
            }
            {
                cudaStream_t s2 = fork_from_recorded(t, 2);
                const float R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
                const float tv[3] = {pose[3], pose[7], pose[11]};
                EF_LAUNCH(t, launch_build_maps(v4, n4, t->height, t->width, t->vmap_g_prev, t->nmap_g_prev, t->tmp_z, R, tv, s2));
                fork_done(t, 2, s2);
            }
        }
    }
    else
    {
        // the five calls, each through its host or device entry
        const float * v4 = static_cast<const float *>(in->vertices_rgba32f);
        const float * n4 = static_cast<const float *>(in->normals_rgba32f);
        const uint8_t * mrgba = static_cast<const uint8_t *>(in->model_rgba8);
        const uint16_t * depth = static_cast<const uint16_t *>(in->depth);
        const uint8_t * rgba = static_cast<const uint8_t *>(in->rgba8);
        rc = host_model ? ef_init_icp_model_host(t, v4, n4, in->depth_cutoff, pose) : ef_init_icp_model(t, v4, n4, in->depth_cutoff, pose);
        if(!rc) rc = host_model ? ef_init_rgb_model_host(t, mrgba) : ef_init_rgb_model(t, mrgba, 0);
        if(!rc) rc = host_sensor ? ef_init_icp_depth_host(t, depth, in->depth_cutoff) : ef_init_icp_depth(t, depth, 0, in->depth_cutoff);
        if(!rc) rc = host_sensor ? ef_init_rgb_host(t, rgba) : ef_init_rgb(t, rgba, 0);
    }
    return rc;
}

EF_API int ef_track_frame_to_model_launch(ef_tracker * t, const ef_frame_inputs * in, const float * pose, int rgb_only, float icp_weight, int pyramid,
                                          int fast_odom, int so3)
{
    EF_ON_DEVICE(t);
    const int rc = frame_build_all(t, in, pose);
    if(rc) return rc;
    const float trans[3] = {pose[3], pose[7], pose[11]};
    const float rot[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
    return ef_get_incremental_transformation_launch(t, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
}

// ------------------------------------------------------------------------------------------------
// k independent sequences per launch (BASELINE.json configs[4]: "k sequences per GPU to show SM fill")
// ------------------------------------------------------------------------------------------------
EF_API int ef_batch_width(void) { return device_track_batch_width(); }

EF_API int ef_track_frames_to_model_batch_launch(ef_tracker * const * ts, int n, const ef_frame_inputs * in, const float * poses, int rgb_only,
                                                 float icp_weight, int pyramid, int fast_odom, int so3)
{
    if(!ts || !in || !poses || n < 2 || n > device_track_batch_width() || n > 8) return EF_ERR_INVALID_ARGUMENT;
    for(int g = 0; g < n; g++)
    {
        if(!ts[g]) return EF_ERR_INVALID_ARGUMENT;
        for(int k = 0; k < g; k++)
            if(ts[k] == ts[g]) return EF_ERR_INVALID_ARGUMENT;
    }
    EF_ON_DEVICE(ts[0]);
    const bool icp = !rgb_only && icp_weight > 0, rgb = rgb_only || icp_weight < 100;
    if(!icp && !rgb) return fail(ts[0], EF_ERR_INVALID_ARGUMENT, "neither ICP nor RGB selected");
    const float * trans[8], * rot[8];
    float tr[8][3], ro[8][9];
    for(int g = 0; g < n; g++)
    {
        ef_tracker * t = ts[g];
        if(t->device != ts[0]->device) return fail(ts[0], EF_ERR_INVALID_ARGUMENT, "the handles of a batch must live on one device");
        if(t->width != ts[0]->width || t->height != ts[0]->height) return fail(ts[0], EF_ERR_INVALID_ARGUMENT, "the handles of a batch must share one image size");
        if(t->solve_mode != EF_SOLVE_DEVICE) return fail(t, EF_ERR_BAD_STATE, "batched tracking needs EF_SOLVE_DEVICE");
        if(t->launch_pending) return fail(t, EF_ERR_BAD_STATE, "a launch is already pending");
    }
    for(int g = 0; g < n; g++)
    {
        ef_tracker * t = ts[g];
        const float * pose = poses + 16 * g;
        int rc = frame_build_all(t, in + g, pose);
        if(rc) return rc;
        if(rgb && !t->deriv_valid && !pyramid)
        {
            rc = compute_derivatives(t);
            if(rc) return rc;
        }
        rc = join_streams(t);
        if(rc) return rc;
        const float tt[3] = {pose[3], pose[7], pose[11]};
        const float rr[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
        memcpy(tr[g], tt, sizeof(tt));
        memcpy(ro[g], rr, sizeof(rr));
        trans[g] = tr[g];
        rot[g] = ro[g];
        memcpy(t->pending.trans, tt, sizeof(tt));
        memcpy(t->pending.rot, rr, sizeof(rr));
        t->pending.rgb_only = rgb_only; t->pending.icp_weight = icp_weight; t->pending.pyramid = pyramid;
        t->pending.fast_odom = fast_odom; t->pending.so3 = so3;
        // the batched kernel runs on the first handle's stream: it waits for the builders of the others ...
        if(g > 0)
        {
            EF_CUDA(t, cudaEventRecord(t->ev_fork, t->stream));
            EF_CUDA(t, cudaStreamWaitEvent(ts[0]->stream, t->ev_fork, 0));
        }
    }
    const int rc = device_track_launch_batch(ts, n, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3, ts[0]->stream);
    if(rc) return rc;
    // ... and whatever the other handles enqueue next on their own streams waits for it
    EF_CUDA(ts[0], cudaEventRecord(ts[0]->ev_fork, ts[0]->stream));
    for(int g = 0; g < n; g++)
    {
        if(g > 0) EF_CUDA(ts[g], cudaStreamWaitEvent(ts[g]->stream, ts[0]->ev_fork, 0));
        if(rgb) ts[g]->deriv_valid = true;
        ts[g]->launch_pending = true;
    }
    return EF_OK;
}

EF_API int ef_track_frames_to_model_batch(ef_tracker * const * ts, int n, const ef_frame_inputs * in, const float * poses, float * trans, float * rot,
                                          int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3, ef_track_stats * stats)
{
    if(!trans || !rot) return EF_ERR_INVALID_ARGUMENT;
    int rc = ef_track_frames_to_model_batch_launch(ts, n, in, poses, rgb_only, icp_weight, pyramid, fast_odom, so3);
    if(rc) return rc;
    for(int g = 0; g < n; g++)
    {
        const int r = ef_get_incremental_transformation_finish(ts[g], trans + 3 * g, rot + 9 * g, stats ? stats + g : nullptr);
        if(r && !rc) rc = r;
    }
    return rc;
}

EF_API int ef_track_frame_to_model(ef_tracker * t, const ef_frame_inputs * in, const float * pose, float * trans, float * rot, int rgb_only,
                                   float icp_weight, int pyramid, int fast_odom, int so3, ef_track_stats * stats)
{
    EF_ON_DEVICE(t);
    const int rc = ef_track_frame_to_model_launch(t, in, pose, rgb_only, icp_weight, pyramid, fast_odom, so3);
    if(rc) return rc;
    return ef_get_incremental_transformation_finish(t, trans, rot, stats);
}

// RGBDOdometry.cpp:605-608
EF_API int ef_get_covariance(ef_tracker * t, double * cov)
{
    if(!t || !cov) return EF_ERR_INVALID_ARGUMENT;
    hm::inverseNN<double, 6>(t->st.last_A, cov);
    return EF_OK;
}

EF_API int ef_tracker_download(ef_tracker * t, const char * name, int level, void * dst, size_t bytes)
{
    EF_ON_DEVICE(t);
    if(!t || !name || !dst || level < 0 || level >= kNumPyrs) return EF_ERR_INVALID_ARGUMENT;
    {
        const int rc_flush = flush_deferred(t);
        if(rc_flush) return rc_flush;
    }
    const size_t n = t->dims[level].n();
    const void * src = nullptr;
    size_t need = 0;
    if(!strcmp(name, "vmap_curr")) { src = t->vmap_curr[level]; need = n * 12; }
    else if(!strcmp(name, "nmap_curr")) { src = t->nmap_curr[level]; need = n * 12; }
    else if(!strcmp(name, "vmap_g_prev")) { src = t->vmap_g_prev[level]; need = n * 12; }
    else if(!strcmp(name, "nmap_g_prev")) { src = t->nmap_g_prev[level]; need = n * 12; }
    else if(!strcmp(name, "last_depth")) { src = t->last_depth[level]; need = n * 4; }
    else if(!strcmp(name, "next_depth")) { src = t->next_depth[level]; need = n * 4; }
    else if(!strcmp(name, "last_image")) { src = t->last_image[level]; need = n; }
    else if(!strcmp(name, "next_image")) { src = t->next_image[level]; need = n; }
    else if(!strcmp(name, "last_next_image")) { src = t->last_next_image[level]; need = n; }
    else if(!strcmp(name, "dIdx")) { src = t->dIdx[level]; need = n * 2; }
    else if(!strcmp(name, "dIdy")) { src = t->dIdy[level]; need = n * 2; }
    else if(!strcmp(name, "depth_tmp")) { src = t->depth_tmp[level]; need = n * 2; }
    else if(!strcmp(name, "filt_depth") && level == 0) { src = t->filt_depth; need = n * 2; }
    else return fail(t, EF_ERR_INVALID_ARGUMENT, "unknown buffer name");
    if(bytes < need) return fail(t, EF_ERR_INVALID_ARGUMENT, "destination too small");
    {
        const int rc = join_streams(t);
        if(rc) return rc;
    }
    EF_CUDA(t, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, t->stream));
    EF_CUDA(t, cudaStreamSynchronize(t->stream));
    return EF_OK;
}

// ------------------------------------------------------------------------------------------------
// Tier 2
// ------------------------------------------------------------------------------------------------
#define EF_OP_RET(call)                          \
    do                                           \
    {                                            \
        if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE; \
        return (int)(call);                      \
    } while(0)

EF_API int ef_op_pyr_down_u16(const uint16_t * s, size_t sp, int r, int c, uint16_t * d, size_t dp, void * st)
{
    EF_OP_RET(launch_pyr_down_u16(s, sp, r, c, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_create_vmap(const uint16_t * depth, size_t dp, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff, float * v,
                             size_t vp, void * st)
{
    EF_OP_RET(launch_create_vmap(depth, dp, rows, cols, fx, fy, cx, cy, cutoff, v, vp, (cudaStream_t)st));
}
EF_API int ef_op_create_nmap(const float * v, size_t vp, int rows, int cols, float * n, size_t np, void * st)
{
    EF_OP_RET(launch_create_nmap(v, vp, rows, cols, n, np, (cudaStream_t)st));
}
EF_API int ef_op_transform_maps(const float * vs, const float * ns, size_t sp, int rows, int cols, const float * R, const float * tv, float * vd,
                                float * nd, size_t dp, void * st)
{
    EF_OP_RET(launch_transform_maps(vs, ns, sp, rows, cols, R, tv, vd, nd, dp, (cudaStream_t)st));
}
EF_API int ef_op_copy_maps(const float * v4, const float * n4, int rows, int cols, float * v, float * n, size_t dp, void * st)
{
    EF_OP_RET(launch_copy_maps(v4, n4, rows, cols, v, n, dp, (cudaStream_t)st));
}
EF_API int ef_op_resize_map(const float * in, size_t ip, int srows, int scols, float * out, size_t op, int normalize, void * st)
{
    EF_OP_RET(launch_resize_map(in, ip, srows, scols, out, op, normalize, (cudaStream_t)st));
}
EF_API int ef_op_vertices_to_depth(const float * v4, int rows, int cols, float cutoff, float * d, size_t dp, void * st)
{
    EF_OP_RET(launch_vertices_to_depth(v4, rows, cols, cutoff, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_pyr_down_gauss_f32(const float * s, size_t sp, int r, int c, float * d, size_t dp, void * st)
{
    EF_OP_RET(launch_pyr_down_gauss_f32(s, sp, r, c, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_pyr_down_gauss_u8(const uint8_t * s, size_t sp, int r, int c, uint8_t * d, size_t dp, void * st)
{
    EF_OP_RET(launch_pyr_down_gauss_u8(s, sp, r, c, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_bgr_to_intensity(const uint8_t * s, size_t sp, int r, int c, uint8_t * d, size_t dp, void * st)
{
    EF_OP_RET(launch_bgr_to_intensity(s, sp, r, c, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_depth_bilateral(const uint16_t * s, size_t sp, int r, int c, float max_depth_m, uint16_t * d, size_t dp, void * st)
{
    EF_OP_RET(launch_depth_bilateral(s, sp, r, c, max_depth_m, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_depth_metric(const uint16_t * s, size_t sp, int r, int c, float max_depth_m, float * d, size_t dp, void * st)
{
    EF_OP_RET(launch_depth_metric(s, sp, r, c, max_depth_m, d, dp, (cudaStream_t)st));
}
EF_API int ef_op_derivative_images(const uint8_t * s, size_t sp, int r, int c, int16_t * dx, int16_t * dy, size_t dp, void * st)
{
    EF_OP_RET(launch_derivative_images(s, sp, r, c, dx, dy, dp, (cudaStream_t)st));
}
EF_API int ef_op_project_point_cloud(const float * depth, size_t dp, int rows, int cols, float fx, float fy, float cx, float cy, int level,
                                     float * cloud, size_t cp, void * st)
{
    const int div = 1 << level; // CameraModel::operator()(level)
    EF_OP_RET(launch_project_points(depth, dp, rows, cols, fx / div, fy / div, cx / div, cy / div, cloud, cp, (cudaStream_t)st));
}

EF_API size_t ef_op_splat_scratch_bytes(int rows, int cols) { return (rows > 0 && cols > 0) ? (size_t)rows * cols * 24 : 0; } // keys + rays

static int splat_predict(const float * surfels, size_t stride_bytes, int count, const float * t_inv, float cx, float cy, float fx, float fy, int rows,
                         int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta, void * keys, uint8_t * image,
                         float * vertex, float * normal, uint16_t * time_out, uint8_t * inst, void * st)
{
    if(count < 0 || (count > 0 && !surfels) || !t_inv || !keys || !vertex || !normal || rows <= 0 || cols <= 0) return EF_ERR_INVALID_ARGUMENT;
    if(stride_bytes < 48 || (stride_bytes % 16) || (reinterpret_cast<uintptr_t>(surfels) % 16) || (reinterpret_cast<uintptr_t>(keys) % 16))
        return EF_ERR_INVALID_ARGUMENT;
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    SplatArgs a;
    a.surfels = surfels; a.stride_bytes = stride_bytes; a.count = count;
    memcpy(a.t_inv, t_inv, sizeof(a.t_inv)); // rows 0..2
    a.cx = cx; a.cy = cy; a.fx = fx; a.fy = fy;
    a.rows = rows; a.cols = cols;
    a.max_depth = max_depth; a.conf_threshold = conf_threshold;
    a.time = time; a.max_time = max_time; a.time_delta = time_delta;
    a.keys = keys; a.image = image; a.vertex = vertex; a.normal = normal; a.time_out = time_out; a.inst = inst;
    return (int)launch_splat_predict(a, (cudaStream_t)st);
}

EF_API int ef_op_splat_predict(const float * surfels, size_t stride_bytes, int count, const float * t_inv, float cx, float cy, float fx, float fy,
                               int rows, int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta, void * keys,
                               uint8_t * image, float * vertex, float * normal, uint16_t * time_out, void * st)
{
    return splat_predict(surfels, stride_bytes, count, t_inv, cx, cy, fx, fy, rows, cols, max_depth, conf_threshold, time, max_time, time_delta, keys,
                         image, vertex, normal, time_out, nullptr, st);
}

EF_API int ef_op_splat_predict_inst(const float * surfels, size_t stride_bytes, int count, const float * t_inv, float cx, float cy, float fx, float fy,
                                    int rows, int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta, void * keys,
                                    uint8_t * image, float * vertex, float * normal, uint16_t * time_out, uint8_t * inst, void * st)
{
    return splat_predict(surfels, stride_bytes, count, t_inv, cx, cy, fx, fy, rows, cols, max_depth, conf_threshold, time, max_time, time_delta, keys,
                         image, vertex, normal, time_out, inst, st);
}

EF_API int ef_op_fill_vertex(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy,
                             int passthrough, float * out, void * st)
{
    if(!predicted || !depth || !out || rows <= 0 || cols <= 0) return EF_ERR_INVALID_ARGUMENT;
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    return (int)launch_fill_vertex(predicted, depth, rows, cols, cx, cy, fx, fy, passthrough, out, (cudaStream_t)st);
}

EF_API int ef_op_fill_normal(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy,
                             int passthrough, float * out, void * st)
{
    if(!predicted || !depth || !out || rows <= 0 || cols <= 0) return EF_ERR_INVALID_ARGUMENT;
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    return (int)launch_fill_normal(predicted, depth, rows, cols, cx, cy, fx, fy, passthrough, out, (cudaStream_t)st);
}

EF_API int ef_op_fill_rgb(const uint8_t * predicted, const uint8_t * raw, int rows, int cols, int passthrough, uint8_t * out, void * st)
{
    if(!predicted || !raw || !out || rows <= 0 || cols <= 0) return EF_ERR_INVALID_ARGUMENT;
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    return (int)launch_fill_rgb(predicted, raw, rows, cols, passthrough, out, (cudaStream_t)st);
}

static int op_fetch(void * scratch, void * host, size_t bytes, cudaStream_t s)
{
    cudaError_t e = cudaMemcpyAsync(host, static_cast<char *>(scratch) + kScratchResultOff, bytes, cudaMemcpyDeviceToHost, s);
    if(e == cudaSuccess) e = cudaStreamSynchronize(s);
    return (int)e;
}

static int op_reset_ticket(void * scratch, cudaStream_t s) { return (int)cudaMemsetAsync(scratch, 0, 128, s); }

EF_API int ef_op_icp_step(const float * Rcurr, const float * tcurr, const float * vc, const float * nc, const float * Rprev_inv, const float * tprev,
                          float fx, float fy, float cx, float cy, const float * vp, const float * np, size_t pitch, float dist_thresh,
                          float angle_thresh, int rows, int cols, void * scratch, float * A, float * b, float * residual, void * st)
{
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    if(!scratch) return EF_ERR_INVALID_ARGUMENT;
    IcpArgs a;
    memcpy(a.Rcurr, Rcurr, 36); memcpy(a.tcurr, tcurr, 12); memcpy(a.Rprev_inv, Rprev_inv, 36); memcpy(a.tprev, tprev, 12);
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.dist_thresh = dist_thresh; a.angle_thresh = angle_thresh;
    a.rows = rows; a.cols = cols;
    a.vmap_curr = vc; a.nmap_curr = nc; a.vmap_g_prev = vp; a.nmap_g_prev = np;
    a.pitch = pitch;
    int rc = op_reset_ticket(scratch, (cudaStream_t)st);
    if(rc) return rc;
    rc = (int)launch_icp_step(a, scratch, (cudaStream_t)st);
    if(rc) return rc;
    float h[29];
    rc = op_fetch(scratch, h, sizeof(h), (cudaStream_t)st);
    if(rc) return rc;
    hm::unpack_se3(h, A, b, residual);
    return EF_OK;
}

EF_API int ef_op_rgb_residual(float min_scale, const int16_t * dIdx, const int16_t * dIdy, size_t d_pitch, const float * last_depth,
                              const float * next_depth, size_t depth_pitch, const uint8_t * last_image, const uint8_t * next_image,
                              size_t image_pitch, void * corres, float max_depth_delta, const float * kt, const float * krkinv, int rows, int cols,
                              void * scratch, int * sigma_sum, int * count, void * st)
{
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    if(!scratch || !corres) return EF_ERR_INVALID_ARGUMENT;
    RgbResArgs a;
    a.min_scale = min_scale; a.max_depth_delta = max_depth_delta;
    memcpy(a.kt, kt, 12); memcpy(a.krkinv, krkinv, 36);
    a.rows = rows; a.cols = cols;
    a.dIdx = dIdx; a.dIdy = dIdy; a.d_pitch = d_pitch;
    a.last_depth = last_depth; a.next_depth = next_depth; a.depth_pitch = depth_pitch;
    a.last_image = last_image; a.next_image = next_image; a.image_pitch = image_pitch;
    a.corres = corres;
    int rc = op_reset_ticket(scratch, (cudaStream_t)st);
    if(rc) return rc;
    rc = (int)launch_rgb_residual(a, scratch, (cudaStream_t)st);
    if(rc) return rc;
    int h[2];
    rc = op_fetch(scratch, h, sizeof(h), (cudaStream_t)st);
    if(rc) return rc;
    *count = h[0];
    *sigma_sum = h[1];
    return EF_OK;
}

EF_API int ef_op_rgb_step(const void * corres, float sigma, const float * cloud, size_t cloud_pitch, float fx, float fy, const int16_t * dIdx,
                          const int16_t * dIdy, size_t d_pitch, float sobel_scale, int rows, int cols, void * scratch, float * A, float * b, void * st)
{
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    if(!scratch) return EF_ERR_INVALID_ARGUMENT;
    RgbStepArgs a;
    a.corres = corres; a.sigma = sigma; a.cloud = cloud; a.cloud_pitch = cloud_pitch;
    a.fx = fx; a.fy = fy; a.dIdx = dIdx; a.dIdy = dIdy; a.d_pitch = d_pitch; a.sobel_scale = sobel_scale;
    a.rows = rows; a.cols = cols;
    int rc = op_reset_ticket(scratch, (cudaStream_t)st);
    if(rc) return rc;
    rc = (int)launch_rgb_step(a, scratch, (cudaStream_t)st);
    if(rc) return rc;
    float h[29];
    rc = op_fetch(scratch, h, sizeof(h), (cudaStream_t)st);
    if(rc) return rc;
    hm::unpack_se3(h, A, b, (float *)nullptr);
    return EF_OK;
}

EF_API int ef_op_so3_step(const uint8_t * last_image, const uint8_t * next_image, size_t image_pitch, const float * image_basis, const float * kinv,
                          const float * krlr, int rows, int cols, void * scratch, float * A, float * b, float * residual, void * st)
{
    if(ef_device_count() <= 0) return EF_ERR_NO_DEVICE;
    if(!scratch) return EF_ERR_INVALID_ARGUMENT;
    So3Args a;
    a.last_image = last_image; a.next_image = next_image; a.image_pitch = image_pitch;
    memcpy(a.image_basis, image_basis, 36); memcpy(a.kinv, kinv, 36); memcpy(a.krlr, krlr, 36);
    a.rows = rows; a.cols = cols;
    int rc = op_reset_ticket(scratch, (cudaStream_t)st);
    if(rc) return rc;
    rc = (int)launch_so3_step(a, scratch, (cudaStream_t)st);
    if(rc) return rc;
    float h[11];
    rc = op_fetch(scratch, h, sizeof(h), (cudaStream_t)st);
    if(rc) return rc;
    hm::unpack_so3(h, A, b, residual);
    return EF_OK;
}
