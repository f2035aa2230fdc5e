/*
 * ef_track.h -- C ABI of the B200-native dense frame-to-model tracker.
 *
 * Drop-in replacement for the CUDA directory of ElasticFusion as shipped in
 * Fancomi2017/InstanceFusion (elasticfusionpublic/Core/src/Cuda: cudafuncs.cu, reduce.cu,
 * containers/) and for the host driver above it (Core/src/Utils/RGBDOdometry.{h,cpp}).
 * File:line citations are relative to elasticfusionpublic/Core/src/.
 *
 * Two tiers:
 *   Tier 1  ef_tracker_* / ef_init_* / ef_get_incremental_transformation
 *           = class RGBDOdometry (Utils/RGBDOdometry.h:31-134), one handle per instance.
 *   Tier 2  ef_op_*  = the 16 free operator functions of Cuda/cudafuncs.cuh:64-177 on raw
 *           (pointer, pitch, rows, cols) images, for callers that own their buffers and for
 *           per-operator parity tests.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t value on a CUDA failure, or a
 *     negative EF_ERR_* code; nothing prints or exits (the reference's cudaSafeCall exits the
 *     process, Cuda/convenience.cuh:64-71).  ef_last_error() gives a message.
 *   - all `d_` pointers are DEVICE pointers borrowed for the duration of the call; `h_` pointers
 *     are host pointers.  Outputs are written to caller-provided host memory.
 *   - a handle is single-threaded; distinct handles share no state (own stream, no globals), so
 *     N handles x M GPUs can be driven from N x M host threads.
 *   - "map3" = 3-plane SoA float image, 3*rows x cols, component c of pixel (x,y) at row
 *     y + c*rows (Utils/RGBDOdometry.cpp:97-101).  "rgba32f" = 4 interleaved floats per pixel
 *     (GL_RGBA32F texel), "rgba8" = 4 interleaved bytes per pixel.
 *   - 3x3 matrices are row-major float[9] (memory image of the reference's mat33,
 *     Cuda/types.cuh:61-73); poses are row-major float[16].
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with
 *     EF_ERR_NO_DEVICE.
 */
#ifndef EF_TRACK_H_
#define EF_TRACK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EF_OK 0
#define EF_ERR_INVALID_ARGUMENT (-1)
#define EF_ERR_NO_DEVICE (-2)
#define EF_ERR_BAD_STATE (-3)
#define EF_ERR_UNSUPPORTED (-4)

#define EF_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------ */
/* Tier 1: tracker handle                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct ef_tracker ef_tracker;

/* public `last*` fields of RGBDOdometry (Utils/RGBDOdometry.h:64-73) + iteration counts */
typedef struct ef_track_stats
{
    float last_icp_error, last_icp_count;
    float last_rgb_error, last_rgb_count;
    float last_so3_error, last_so3_count;
    double last_A[36]; /* row-major 6x6 */
    double last_b[6];
    int so3_iterations;    /* so3Step calls made (<= 10) */
    int se3_iterations[3]; /* Gauss-Newton iterations run per pyramid level */
} ef_track_stats;

/* ef_tracker_set_option keys */
#define EF_OPT_SOLVE_MODE 1      /* EF_SOLVE_DEVICE (default wherever the image fits the persistent kernel, i.e. up to ~1920x1080 on a
                                    B200) | EF_SOLVE_HOST (the fallback for larger images) */
#define EF_OPT_USE_GRAPH 2       /* 0/1 (host-solve mode): the launches of one Gauss-Newton iteration (computeRgbResidual, rgbStep, icpStep,
                                    parameter upload, result copies) are captured once per pyramid level and replayed as ONE CUDA
                                    graph launch + ONE synchronisation per iteration; results identical to 0.  Device-solve mode
                                    needs no graph: its whole iteration loop is one kernel */
#define EF_OPT_FUSED_BUILD 3     /* 0/1: fused pyramid builders (default 1 when both image sides are multiples of 4, which they need)
                                    instead of one kernel per operator (any size) */
#define EF_OPT_PROFILE 4         /* 0/1: bracket the solve of every getIncrementalTransformation with CUDA events */
#define EF_OPT_GRID_CTAS 5       /* device mode: CTAs (= SMs) the persistent tracker kernel occupies; 0 = all.  Lets k handles
                                    track k independent sequences concurrently on disjoint SMs of one GPU.  A handle on a subset
                                    runs the symmetric build of the kernel (every CTA a worker, each solving redundantly: one L2
                                    hop per iteration instead of two); the order of its float additions, and so its last bits,
                                    depend on the number of CTAs */

#define EF_OPT_AUX_STREAMS 6     /* 0/1 (default 1): the pyramid builders that do not depend on each other (current-frame depth
                                    pyramid, model RGB-D pyramid) run on two internal streams beside the handle's stream and
                                    are joined before the solve; 0 = everything in order on the handle's stream */

#define EF_OPT_FRAME_BUILD 7     /* single-call entry ef_track_frame_to_model: 0 = the builders of the five init* calls (chained
                                    kernels on the internal streams), 1 = every pyramid of the frame from ONE kernel launch when
                                    the inputs are device buffers, 2 (default) = also for host inputs (five copies, then the one
                                    launch) */

#define EF_OPT_DEFER_BUILD 8     /* 0/1 (default 0): ef_init_icp_model / ef_init_rgb_model / ef_init_icp_depth / ef_init_rgb with DEVICE
                                    pointers only record their arguments; the pyramids are built when they are first needed
                                    (ef_get_incremental_transformation, ef_tracker_download, ...) -- all four together from ONE kernel
                                    launch, like ef_track_frame_to_model.  The caller promises that the buffers stay valid and
                                    unchanged until then (the reference borrows its textures only for the duration of an init call) */

#define EF_OPT_HOST_FUSED 9      /* 0/1 (default 1; host-solve mode, image width a multiple of 16, sides < 4096): a Gauss-Newton iteration
                                    is two launches and no copy -- computeRgbResidual writing 8-byte correspondences, then icpStep +
                                    rgbStep with one fixed-order reduction storing all sums straight into mapped pinned memory, which
                                    the host polls.  0 = one launch, one download and one synchronisation per operator, exactly the
                                    reference's control flow (or EF_OPT_USE_GRAPH's replay of it) */

#define EF_SOLVE_HOST 0   /* one step kernel per operator call, 6x6 LDLT + pose update in double on the host,
                             exactly the reference's control flow (RGBDOdometry.cpp:405-585) */
#define EF_SOLVE_DEVICE 1 /* one persistent cooperative kernel runs the SO(3) loop and all Gauss-Newton
                             iterations; the same double-precision LDLT runs in a single device thread */

/* RGBDOdometry::RGBDOdometry (Utils/RGBDOdometry.cpp:21-111).  Any width, height >= 16 (level i is (width >> i) x (height >> i)
 * like the reference's).  `stream` is a cudaStream_t or NULL (the handle then creates its own non-blocking stream).
 * The handle lives on the CUDA device that is current at this call; every later call on the handle makes that device
 * current for its duration and restores the caller's.
 * STREAM CONTRACT: device inputs (`d_` pointers, cudaArrays) are consumed on the handle's stream.  Work that produces them
 * on another stream must be ordered before the call by the caller: record an event there and pass it to
 * ef_tracker_wait_event (or hand the stream to ef_tracker_wait_stream), or synchronise.  (The reference has no such issue: everything runs on the legacy default stream.) */
int ef_tracker_create(int width, int height, float cx, float cy, float fx, float fy, float dist_thresh, float angle_thresh,
                      void * stream, ef_tracker ** out);
int ef_tracker_destroy(ef_tracker * t);
int ef_tracker_set_option(ef_tracker * t, int key, int value);
int ef_tracker_get_option(ef_tracker * t, int key, int * value);
const char * ef_last_error(const ef_tracker * t);
void * ef_tracker_stream(ef_tracker * t);
int ef_tracker_synchronize(ef_tracker * t);
/* the handle's stream waits for `cuda_event` (a cudaEvent_t recorded on the stream that produces the next call's device inputs) */
int ef_tracker_wait_event(ef_tracker * t, void * cuda_event);
/* the same from the producer's stream itself (a cudaStream_t; NULL = the legacy default stream): nothing happens when that stream is
 * idle or is the handle's own stream, otherwise the handle's stream waits for everything enqueued there so far */
int ef_tracker_wait_stream(ef_tracker * t, void * cuda_stream);

/* default thresholds of the reference constructor (Utils/RGBDOdometry.h:38-39) */
float ef_default_dist_thresh(void);
float ef_default_angle_thresh(void);

/* initICP(GPUTexture * filteredDepth, depthCutoff)                 RGBDOdometry.cpp:118-142
 * d_depth: uint16 millimetres, rows of `pitch_bytes` bytes (0 = dense). */
int ef_init_icp_depth(ef_tracker * t, const uint16_t * d_depth, size_t pitch_bytes, float depth_cutoff);
/* filterDepth + initICP in one call (ElasticFusion.cpp:309 + :348): RAW sensor depth in, bilateral filter with the range gate
 * [300 mm, max_depth_m] on the way (into a buffer the handle owns), then as ef_init_icp_depth */
int ef_init_icp_depth_raw(ef_tracker * t, const uint16_t * d_raw_depth, size_t pitch_bytes, float max_depth_m, float depth_cutoff);
int ef_init_icp_depth_raw_host(ef_tracker * t, const uint16_t * h_raw_depth, float max_depth_m, float depth_cutoff);
/* initICP(GPUTexture * predictedVertices, GPUTexture * predictedNormals, depthCutoff)   :144-167 */
int ef_init_icp_maps(ef_tracker * t, const float * d_vertices_rgba32f, const float * d_normals_rgba32f, float depth_cutoff);
/* initICPModel(predictedVertices, predictedNormals, depthCutoff, modelPose)             :169-206 */
int ef_init_icp_model(ef_tracker * t, const float * d_vertices_rgba32f, const float * d_normals_rgba32f, float depth_cutoff,
                      const float * h_pose16);
/* initRGB / initRGBModel / initFirstRGB (GPUTexture * rgb)                  :243-247, :237-241, :249-265
 * NOTE (reference call-order contract, ElasticFusion.cpp:342): an initICP* call must precede
 * initRGB*; the float depth pyramid is taken from the vertex map most recently passed to
 * ef_init_icp_maps / ef_init_icp_model (the reference's vmaps_tmp reuse, RGBDOdometry.cpp:212). */
int ef_init_rgb(ef_tracker * t, const uint8_t * d_rgba8, size_t pitch_bytes);
int ef_init_rgb_model(ef_tracker * t, const uint8_t * d_rgba8, size_t pitch_bytes);
int ef_init_first_rgb(ef_tracker * t, const uint8_t * d_rgba8, size_t pitch_bytes);

/* cudaArray_t variants for CUDA<->GL interop callers (GPUTexture::cudaRes mapped with
 * cudaGraphicsSubResourceGetMappedArray, RGBDOdometry.cpp:120-128): the array is copied into a
 * handle-owned linear staging buffer on the handle's stream, then the pointer variant runs. */
int ef_init_icp_depth_array(ef_tracker * t, void * cuda_array, float depth_cutoff);
int ef_init_icp_maps_array(ef_tracker * t, void * vertices_array, void * normals_array, float depth_cutoff);
int ef_init_icp_model_array(ef_tracker * t, void * vertices_array, void * normals_array, float depth_cutoff,
                            const float * h_pose16);
int ef_init_rgb_array(ef_tracker * t, void * rgba_array);
int ef_init_rgb_model_array(ef_tracker * t, void * rgba_array);
int ef_init_first_rgb_array(ef_tracker * t, void * rgba_array);

/* host-buffer variants: asynchronous H2D into the handle's staging buffers (pinned host memory makes
 * them truly asynchronous), then the pointer variant.  Dense rows. */
int ef_init_icp_depth_host(ef_tracker * t, const uint16_t * h_depth, float depth_cutoff);
int ef_init_icp_maps_host(ef_tracker * t, const float * h_vertices_rgba32f, const float * h_normals_rgba32f, float depth_cutoff);
int ef_init_icp_model_host(ef_tracker * t, const float * h_vertices_rgba32f, const float * h_normals_rgba32f, float depth_cutoff,
                           const float * h_pose16);
int ef_init_rgb_host(ef_tracker * t, const uint8_t * h_rgba8);
int ef_init_rgb_model_host(ef_tracker * t, const uint8_t * h_rgba8);
int ef_init_first_rgb_host(ef_tracker * t, const uint8_t * h_rgba8);

/* getIncrementalTransformation(trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3)   :267-603
 * trans[3] / rot[9] (row-major) are in/out exactly like the reference's Eigen references.
 * Blocks until the result is on the host (the reference blocks on every step). `stats` may be NULL. */
int ef_get_incremental_transformation(ef_tracker * t, float * trans3, float * rot9, int rgb_only, float icp_weight,
                                      int pyramid, int fast_odom, int so3, ef_track_stats * stats);
/* asynchronous split of the call above: _launch enqueues all work and returns, _finish waits for it
 * and writes the outputs.  Exactly one _finish per _launch. */
int ef_get_incremental_transformation_launch(ef_tracker * t, const float * trans3, const float * rot9, int rgb_only,
                                             float icp_weight, int pyramid, int fast_odom, int so3);
int ef_get_incremental_transformation_finish(ef_tracker * t, float * trans3, float * rot9, ef_track_stats * stats);

/* The frameToModel call sequence of ElasticFusion::processFrame (ElasticFusion.cpp:343-368) as one call:
 *   initICPModel(vertices, normals, depth_cutoff, pose) ; initRGBModel(model_rgba8) ; initICP(depth, depth_cutoff) ;
 *   initRGB(rgba8) ; getIncrementalTransformation(trans = pose.t, rot = pose.R, ...)
 * `on_host`: 0 = all five images are device pointers, 1 = all are host pointers (dense rows, ideally pinned),
 * 2 = the SENSOR frame (depth, rgba8) is on the host and the MODEL maps (vertices, normals, model_rgba8) are device
 * pointers -- the production data flow: the model prediction lives on the GPU (GL textures in the reference,
 * ef_op_splat_predict here) and only the 1.8 MB sensor frame crosses PCIe.  Same results as the five separate calls. */
typedef struct ef_frame_inputs
{
    const float * vertices_rgba32f;
    const float * normals_rgba32f;
    const uint8_t * model_rgba8;
    const uint16_t * depth;
    const uint8_t * rgba8;
    float depth_cutoff;
    int on_host;
} ef_frame_inputs;
int ef_track_frame_to_model_launch(ef_tracker * t, const ef_frame_inputs * in, const float * h_pose16, int rgb_only, float icp_weight,
                                   int pyramid, int fast_odom, int so3);
/* = _launch + ef_get_incremental_transformation_finish */
int ef_track_frame_to_model(ef_tracker * t, const ef_frame_inputs * in, const float * h_pose16, float * trans3, float * rot9, int rgb_only,
                            float icp_weight, int pyramid, int fast_odom, int so3, ef_track_stats * stats);

/* k INDEPENDENT sequences per kernel launch (BASELINE.json configs[4]: "k sequences per GPU to show SM fill").
 * handles[0 .. n): n = 2 .. ef_batch_width() (4) distinct handles of one image size on one device, EF_SOLVE_DEVICE; inputs[g] and
 * poses16[16 * g ..] are what ef_track_frame_to_model takes for handle g.  Every handle's pyramids are built as usual, then
 * ONE persistent tracker kernel runs the Gauss-Newton solves of all n frames: sequence g has its own solver CTA and the other
 * CTAs take the sequences in turn, one iteration each, so the serial gather -> solve -> publish chain of one sequence (~3 us
 * per iteration, during which a single launch leaves the SMs idle) is hidden behind the pixels of the others.  Per handle
 * the results are bit-identical to ef_track_frame_to_model on a handle configured with EF_OPT_GRID_CTAS = SMs - n (the same
 * number of worker CTAs) and within the pose tolerance of the default launch.  _launch returns once everything is enqueued;
 * finish each handle with ef_get_incremental_transformation_finish (any order) or use the blocking form, which writes
 * trans[3 * g ..], rot[9 * g ..] and stats[g] (may be NULL).  EF_ERR_UNSUPPORTED: the image is too large for a sequence's share
 * of the workers' shared memory (about 800 x 600 for two sequences on a B200).  The first handle's EF_OPT_GRID_CTAS sizes the
 * launch.  (Environment EF_BATCH_MODE=groups: the older build, exactly two sequences as two thread groups per CTA, the
 * default single launch's bits.) */
int ef_batch_width(void);
int ef_track_frames_to_model_batch_launch(ef_tracker * const * handles, int n, const ef_frame_inputs * inputs, const float * h_poses16, int rgb_only,
                                          float icp_weight, int pyramid, int fast_odom, int so3);
int ef_track_frames_to_model_batch(ef_tracker * const * handles, int n, const ef_frame_inputs * inputs, const float * h_poses16, float * trans3n,
                                   float * rot9n, int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3, ef_track_stats * stats);

/* getCovariance(): inverse of lastA                                                    :605-608 */
int ef_get_covariance(ef_tracker * t, double * cov36);

/* Test/diagnostic access to the handle's pyramids (dense rows): name in
 * {"vmap_curr","nmap_curr","vmap_g_prev","nmap_g_prev","last_depth","next_depth","last_image",
 *  "next_image","last_next_image","dIdx","dIdy","depth_tmp"}, level 0..2.  Synchronises the stream.
 * Note dIdx/dIdy are only (re)computed inside ef_get_incremental_transformation when RGB is used. */
int ef_tracker_download(ef_tracker * t, const char * name, int level, void * h_dst, size_t bytes);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
long long ef_tracker_launch_count(const ef_tracker * t);
/* EF_OPT_PROFILE: device time (CUDA events on the handle's stream) spent in the solve part of
 * getIncrementalTransformation -- the persistent tracker kernel in EF_SOLVE_DEVICE, the whole step loop in
 * EF_SOLVE_HOST -- accumulated over `calls` calls since the last read; reading resets the accumulator. */
int ef_tracker_profile(ef_tracker * t, double * solve_ms_total, long long * calls);
/* The reference's Stopwatch keys on this path (Utils/RGBDOdometry.cpp:333/346 "so3Step", :441/458 "computeRgbResidual", :493/512
 * "icpStep", :523/538 "rgbStep"; Utils/Stopwatch.h:59-82): wall-clock milliseconds around the blocking operator call, the LAST
 * call of each operator winning like the reference's TICK / TOCK pairs (GPUTest.cpp:283-286 reads exactly these), plus their sums
 * over the last getIncrementalTransformation.  Filled in EF_SOLVE_HOST, where the operators are separate blocking calls; with
 * EF_OPT_USE_GRAPH one replayed graph holds all operators of an iteration and its time goes to iteration_ms; in EF_SOLVE_DEVICE
 * the whole solve is one kernel: only call_ms is set (per-level device times: ef_tracker_trace, whole-solve device time:
 * ef_tracker_profile).  include/compat/RGBDOdometry.h forwards them to Stopwatch when EF_COMPAT_WITH_STOPWATCH is defined. */
typedef struct ef_stage_times
{
    float so3_step_ms, rgb_residual_ms, icp_step_ms, rgb_step_ms;                 /* last call of each operator */
    float so3_step_sum_ms, rgb_residual_sum_ms, icp_step_sum_ms, rgb_step_sum_ms; /* over the last getIncrementalTransformation */
    float iteration_ms, iteration_sum_ms;                                         /* EF_OPT_USE_GRAPH: one replayed iteration */
    float call_ms;                                                                /* the whole getIncrementalTransformation, launch to result */
    int solve_mode;                                                               /* EF_SOLVE_HOST | EF_SOLVE_DEVICE of that call */
} ef_stage_times;
int ef_tracker_stage_times(ef_tracker * t, ef_stage_times * out);

/* Per-iteration trace of the persistent tracker kernel, for handles created with EF_TRACK_TIMING=1 in the environment (the
 * kernel then stamps clock64 at its phase boundaries; EF_TRACK_TIMING_PRINT=1 also prints the phase table when the handle
 * is destroyed): cycles32[i] = SM clock of the solver CTA at the top of Gauss-Newton iteration i (in launch order: level 2
 * first), relative to the kernel's start and summed over `calls` calls; cycles32[31] = the kernel's end.  EF_ERR_BAD_STATE
 * without EF_TRACK_TIMING.  bench.py derives the per-pyramid-level share of a launch from it. */
int ef_tracker_trace(ef_tracker * t, double * cycles32, long long * calls);

/* ------------------------------------------------------------------------------------------ */
/* Tier 2: per-operator entry points (Cuda/cudafuncs.cuh:64-177).  All asynchronous on `stream`  */
/* (cudaStream_t, NULL = legacy default stream) unless they return host results, in which case   */
/* they synchronise `stream` before returning.  Pitches are in BYTES (0 = dense).                */
/* ------------------------------------------------------------------------------------------ */
/* pyrDown                     cudafuncs.cu:57-107   u16 src rows x cols -> dst rows/2 x cols/2 */
int ef_op_pyr_down_u16(const uint16_t * d_src, size_t src_pitch, int src_rows, int src_cols, uint16_t * d_dst, size_t dst_pitch,
                       void * stream);
/* createVMap                  cudafuncs.cu:109-149  fx..cy already scaled to the level */
int ef_op_create_vmap(const uint16_t * d_depth, size_t depth_pitch, int rows, int cols, float fx, float fy, float cx, float cy,
                      float depth_cutoff, float * d_vmap3, size_t vmap_pitch, void * stream);
/* createNMap                  cudafuncs.cu:151-204 */
int ef_op_create_nmap(const float * d_vmap3, size_t vmap_pitch, int rows, int cols, float * d_nmap3, size_t nmap_pitch, void * stream);
/* tranformMaps                cudafuncs.cu:206-268  (dst may alias src) */
int ef_op_transform_maps(const float * d_vsrc3, const float * d_nsrc3, size_t src_pitch, int rows, int cols, const float * h_R9,
                         const float * h_t3, float * d_vdst3, float * d_ndst3, size_t dst_pitch, void * stream);
/* copyMaps                    cudafuncs.cu:270-330 */
int ef_op_copy_maps(const float * d_v_rgba32f, const float * d_n_rgba32f, int rows, int cols, float * d_vmap3, float * d_nmap3,
                    size_t dst_pitch, void * stream);
/* resizeVMap / resizeNMap     cudafuncs.cu:365-444  normalize = 0 / 1 */
int ef_op_resize_map(const float * d_in3, size_t in_pitch, int src_rows, int src_cols, float * d_out3, size_t out_pitch,
                     int normalize, void * stream);
/* verticesToDepth             cudafuncs.cu:526-546 */
int ef_op_vertices_to_depth(const float * d_v_rgba32f, int rows, int cols, float cutoff, float * d_dst, size_t dst_pitch, void * stream);
/* pyrDownGaussF               cudafuncs.cu:332-363, 446-468 */
int ef_op_pyr_down_gauss_f32(const float * d_src, size_t src_pitch, int src_rows, int src_cols, float * d_dst, size_t dst_pitch,
                             void * stream);
/* pyrDownUcharGauss           cudafuncs.cu:470-524 */
int ef_op_pyr_down_gauss_u8(const uint8_t * d_src, size_t src_pitch, int src_rows, int src_cols, uint8_t * d_dst, size_t dst_pitch,
                            void * stream);
/* imageBGRToIntensity         cudafuncs.cu:548-577  (linear rgba8 source instead of a texture reference) */
int ef_op_bgr_to_intensity(const uint8_t * d_rgba8, size_t src_pitch, int rows, int cols, uint8_t * d_dst, size_t dst_pitch,
                           void * stream);
/* computeDerivativeImages     cudafuncs.cu:580-639 */
int ef_op_derivative_images(const uint8_t * d_src, size_t src_pitch, int rows, int cols, int16_t * d_dx, int16_t * d_dy,
                            size_t d_pitch, void * stream);
/* projectToPointCloud         cudafuncs.cu:641-674  fx..cy = LEVEL-0 intrinsics, divided by 2^level inside;
 * cloud = rows x cols float3 */
int ef_op_project_point_cloud(const float * d_depth, size_t depth_pitch, int rows, int cols, float fx, float fy, float cx, float cy,
                              int level, float * d_cloud3, size_t cloud_pitch, void * stream);

/* ---- the step before the tracker, GLSL in the reference ---- */
/* ElasticFusion::filterDepth   ElasticFusion.cpp:775-784, Shaders/depth_bilateral.frag:30-76: 13x13 bilateral filter of the raw
 * depth in millimetres; pixels outside [300 mm, max_depth_m] become 0 */
int ef_op_depth_bilateral(const uint16_t * d_src, size_t src_pitch, int rows, int cols, float max_depth_m, uint16_t * d_dst,
                          size_t dst_pitch, void * stream);
/* ElasticFusion::metriciseDepth ElasticFusion.cpp:765-773, Shaders/depth_metric.frag:28-40: millimetres -> metres, same gate */
int ef_op_depth_metric(const uint16_t * d_src, size_t src_pitch, int rows, int cols, float max_depth_m, float * d_dst, size_t dst_pitch,
                       void * stream);

/* ---- the step around the tracker, OpenGL in the reference: the model prediction whose outputs are
 * the `vertices_rgba32f / normals_rgba32f / model_rgba8` inputs of ef_init_icp_model / ef_init_rgb_model ---- */
/* IndexMap::combinedPredict   IndexMap.cpp:468-575, Shaders/splat.vert:19-88, Shaders/combo_splat.frag:19-67: every surfel
 * (float4 position|confidence, float4 colour|instance|initTime|time, float4 normal|radius at d_surfels + i * stride_bytes;
 * stride 48 in ElasticFusion, Vertex::SIZE = 256 in InstanceFusion) is drawn as a point sprite under a GL_LESS depth test.
 * h_t_inv16 = inverse of the prediction pose, row-major.  d_keys: ef_op_splat_scratch_bytes(rows, cols) bytes of device
 * scratch.  Outputs rows x cols, zero where nothing was drawn: d_image_rgba8 and d_time_u16 may be null. */
size_t ef_op_splat_scratch_bytes(int rows, int cols);
int ef_op_splat_predict(const float * d_surfels, size_t stride_bytes, int count, const float * h_t_inv16, float cx, float cy, float fx,
                        float fy, int rows, int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta,
                        void * d_keys, uint8_t * d_image_rgba8, float * d_vertex_rgba32f, float * d_normal_rgba32f,
                        uint16_t * d_time_u16, void * stream);
/* The same with InstanceFusion's fifth render target (combo_splat.frag:29, :54: inst = decodeColor(colTime.y)): d_inst_rgba8 receives
 * the instance colour of the winning surfel (second float of its colour vector), zero where nothing was drawn; may be null.
 * (Degenerate fragments: a NaN intersection is discarded here; the shader keeps it and leaves it to the depth test, whose result
 * for a NaN depth is implementation-defined in GL.) */
int ef_op_splat_predict_inst(const float * d_surfels, size_t stride_bytes, int count, const float * h_t_inv16, float cx, float cy, float fx,
                             float fy, int rows, int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta,
                             void * d_keys, uint8_t * d_image_rgba8, float * d_vertex_rgba32f, float * d_normal_rgba32f,
                             uint16_t * d_time_u16, uint8_t * d_inst_rgba8, void * stream);
/* FillIn::vertex / normal / image   FillIn.cpp, Shaders/fill_vertex.frag:19-53, fill_normal.frag:19-55, fill_rgb.frag:19-37
 * (ElasticFusion.cpp:756-760): holes of the predicted maps (vertex z == 0, black colour) are patched from the current raw
 * depth (millimetres) / colour image; passthrough = 1 takes every pixel from the current frame */
int ef_op_fill_vertex(const float * d_predicted_rgba32f, const uint16_t * d_depth_mm, int rows, int cols, float cx, float cy, float fx,
                      float fy, int passthrough, float * d_out_rgba32f, void * stream);
int ef_op_fill_normal(const float * d_predicted_rgba32f, const uint16_t * d_depth_mm, int rows, int cols, float cx, float cy, float fx,
                      float fy, int passthrough, float * d_out_rgba32f, void * stream);
int ef_op_fill_rgb(const uint8_t * d_predicted_rgba8, const uint8_t * d_rgba8, int rows, int cols, int passthrough, uint8_t * d_out_rgba8,
                   void * stream);

/* icpStep                     reduce.cu:257-490.  Host outputs: A 6x6 row-major, b[6], residual[2] =
 * {sum r^2, inlier count}.  `d_scratch` >= ef_op_scratch_bytes() bytes of device memory. */
int ef_op_icp_step(const float * h_Rcurr9, const float * h_tcurr3, const float * d_vmap_curr3, const float * d_nmap_curr3,
                   const float * h_Rprev_inv9, const float * h_tprev3, float fx, float fy, float cx, float cy,
                   const float * d_vmap_g_prev3, const float * d_nmap_g_prev3, size_t map_pitch, float dist_thresh,
                   float angle_thresh, int rows, int cols, void * d_scratch, float * h_A36, float * h_b6, float * h_residual2,
                   void * stream);
/* computeRgbResidual          reduce.cu:739-936.  d_corres = rows*cols 16-byte DataTerm records
 * (Cuda/types.cuh:75-81), linear (the reference indexes it linearly, reduce.cu:515/839). */
int ef_op_rgb_residual(float min_scale, const int16_t * d_dIdx, const int16_t * d_dIdy, size_t d_pitch, const float * d_last_depth,
                       const float * d_next_depth, size_t depth_pitch, const uint8_t * d_last_image, const uint8_t * d_next_image,
                       size_t image_pitch, void * d_corres, float max_depth_delta, const float * h_kt3, const float * h_krkinv9,
                       int rows, int cols, void * d_scratch, int * h_sigma_sum, int * h_count, void * stream);
/* rgbStep                     reduce.cu:494-678 */
int ef_op_rgb_step(const void * d_corres, float sigma, const float * d_cloud3, size_t cloud_pitch, float fx, float fy,
                   const int16_t * d_dIdx, const int16_t * d_dIdy, size_t d_pitch, float sobel_scale, int rows, int cols,
                   void * d_scratch, float * h_A36, float * h_b6, void * stream);
/* so3Step                     reduce.cu:938-1141 */
int ef_op_so3_step(const uint8_t * d_last_image, const uint8_t * d_next_image, size_t image_pitch, const float * h_image_basis9,
                   const float * h_kinv9, const float * h_krlr9, int rows, int cols, void * d_scratch, float * h_A9, float * h_b3,
                   float * h_residual2, void * stream);
size_t ef_op_scratch_bytes(void);

int ef_abi_version(void);
int ef_device_count(void);

#ifdef __cplusplus
}
#endif

#endif /* EF_TRACK_H_ */
