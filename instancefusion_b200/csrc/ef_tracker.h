// ef_tracker.h -- the tracker handle behind the C ABI (include/ef_track.h): device memory arena,
// pyramids, staging buffers, pinned result block.  Mirrors the members of class RGBDOdometry
// (elasticfusionpublic/Core/src/Utils/RGBDOdometry.h:75-133) with dense rows instead of
// cudaMallocPitch'ed DeviceArray2D.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <chrono>
#include <string>

#include "../../include/ef_track.h"

namespace ef
{

constexpr int kNumPyrs = 3; // RGBDOdometry.h:107

struct LevelDims
{
    int rows, cols;
    size_t n() const { return (size_t)rows * cols; }
};

} // namespace ef

struct ef_tracker
{
    int width, height;
    float cx, cy, fx, fy;
    float dist_thresh, angle_thresh;
    float sobel_scale, max_depth_delta_rgb, max_depth_rgb; // RGBDOdometry.cpp:34-37
    float min_grad[ef::kNumPyrs];                          // :107-110
    ef::LevelDims dims[ef::kNumPyrs];

    int device;
    int num_sms;
    cudaStream_t stream;
    bool own_stream;

    int solve_mode;  // EF_SOLVE_HOST | EF_SOLVE_DEVICE
    int use_graph;
    int fused_build;
    bool vector_ok;  // both image sides are multiples of 4: the fused (vectorised) builders apply
    int grid_ctas;   // EF_OPT_GRID_CTAS (0 = every SM)
    int aux_streams; // EF_OPT_AUX_STREAMS
    int frame_build; // EF_OPT_FRAME_BUILD
    int defer_build; // EF_OPT_DEFER_BUILD
    int host_fused;  // EF_OPT_HOST_FUSED
    unsigned iter_seq; // sequence number of the last fused host-mode iteration (ef_iter_fused.cu)
    struct
    {
        unsigned have;            // 1 model maps | 2 model colour | 4 depth | 8 colour
        const float * v, * n;
        float pose[16];
        const uint8_t * model_rgba, * rgba;
        const uint16_t * depth;
        float depth_cutoff;
    } deferred;

    // internal fork/join streams for builders that are independent of each other (ef_api.cu: fork_stream / join_streams)
    cudaStream_t aux[3];            // 0: current-frame depth chain, 1: model RGB-D chain, 2: model maps (single-call entry)
    cudaEvent_t ev_fork, ev_join[3];
    bool aux_dirty[3];

    // one device arena, sliced
    void * arena;
    size_t arena_bytes;

    uint16_t * depth_tmp[ef::kNumPyrs];
    uint16_t * filt_depth; // bilateral-filtered raw depth (ef_init_icp_depth_raw; textures[DEPTH_FILTERED] of ElasticFusion.cpp:188)
    float * tmp_z; // z channel of the vertex map last given to init_icp_maps/init_icp_model (stands for vmaps_tmp)
    float * vmap_curr[ef::kNumPyrs], * nmap_curr[ef::kNumPyrs];
    float * vmap_g_prev[ef::kNumPyrs], * nmap_g_prev[ef::kNumPyrs];
    float * last_depth[ef::kNumPyrs], * next_depth[ef::kNumPyrs];
    uint8_t * last_image[ef::kNumPyrs], * next_image[ef::kNumPyrs], * last_next_image[ef::kNumPyrs];
    int16_t * dIdx[ef::kNumPyrs], * dIdy[ef::kNumPyrs];
    void * corres[ef::kNumPyrs]; // 16-byte DataTerm records (host-solve path)
    float * cloud[ef::kNumPyrs]; // float3 point clouds (host-solve path)
    void * scratch;              // reduction scratch (ef_kernels.h layout)
    void * scratch2;             // a second one: the fused host-mode iteration runs two reducing kernels back to back

    // staging for the _host and _array entry points
    uint16_t * stage_depth;
    uint8_t * stage_rgba, * stage_rgba_model; // two: the model image is read on an internal stream while the next copy runs
    float * stage_v, * stage_n;
    int stage_reader[5]; // internal stream whose builder last read each staging buffer (-1: the handle's stream), ef_api.cu stage_guard
    // sensor frame from the host with the model maps on the device (ef_frame_inputs.on_host == 2, the production data flow): the two
    // copies run on their own stream, beside whatever is already enqueued on the handle's stream (the caller's model prediction)
    cudaStream_t copy_stream;
    cudaEvent_t ev_stage_free, ev_stage_ready; // the last builder that read stage_depth / stage_rgba is done | the copies have landed
    bool stage_free_valid;                     // ev_stage_free covers the last use of the two staging buffers

    // persistent-kernel state (EF_SOLVE_DEVICE)
    int track_variant;  // threads per CTA of the tracker-kernel build this handle uses (ef_track_dispatch.cu)
    void * track_state; // device TrackState
    void * h_track_out; // pinned TrackOutput
    bool launch_pending;
    struct
    {
        float trans[3], rot[9];
        int rgb_only, pyramid, fast_odom, so3;
        float icp_weight;
    } pending;

    float * h_result; // pinned, ef::kHostResultFloats floats: [0, 64) one fetched result; graph replay: see ef_api.cu

    // EF_OPT_USE_GRAPH (host-solve mode): the launches of one Gauss-Newton iteration of a level as a replayed CUDA graph
    // (two sets: the SO(3) pre-alignment swaps the nextImage / lastNextImage pyramids after every call, RGBDOdometry.cpp:593-599,
    //  so the buffers a graph was captured with come back every other call)
    cudaGraphExec_t iter_graph[2][ef::kNumPyrs];
    int iter_graph_key[2][ef::kNumPyrs]; // which operators the graph holds (0 = none built)
    int image_parity;                    // number of image swaps so far, mod 2
    void * graph_scratch;             // three reduction scratch blocks + the device copy of ef::IterParams
    bool deriv_valid; // dIdx/dIdy match next_image

    // EF_OPT_PROFILE
    int profile;
    cudaEvent_t ev_begin, ev_end;
    bool ev_pending;
    double prof_ms;
    long long prof_calls;

    ef_track_stats st;
    ef_stage_times stages;                            // Stopwatch-compatible timings (ef_tracker_stage_times)
    std::chrono::steady_clock::time_point call_begin; // start of the pending getIncrementalTransformation
    std::string err;
    long long launches;
};

namespace ef
{
// EF_SOLVE_DEVICE path (ef_track_kernel.cu)
int device_track_init(ef_tracker * t);
int device_track_configure(ef_tracker * t, int grid_ctas);
int device_track_trace(ef_tracker * t, double * out32, long long * calls);
bool device_track_supported(const ef_tracker * t); // the image fits the persistent kernel's shared-memory candidate store
void device_track_destroy(ef_tracker * t);
int device_track_launch(ef_tracker * t, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom,
                        int so3);
int device_track_finish(ef_tracker * t, float * trans, float * rot);
// the batched build of the tracker kernel: device_track_batch_width() sequences (handles) per launch
int device_track_batch_width();
int device_track_launch_batch(ef_tracker * const * ts, int n, const float * const * trans, const float * const * rot, int rgb_only, float icp_weight,
                              int pyramid, int fast_odom, int so3, cudaStream_t stream);
} // namespace ef
