// ef_kernels.h -- internal launcher declarations shared by the C-ABI layer (ef_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ef
{

// ---- Tier-2 image operators (ef_ops_image.cu); pitches in bytes, 0 = dense ----
cudaError_t launch_pyr_down_u16(const uint16_t * src, size_t sp, int srows, int scols, uint16_t * dst, size_t dp, cudaStream_t s);
cudaError_t launch_create_vmap(const uint16_t * depth, size_t dp, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff,
                               float * vmap, size_t vp, cudaStream_t s);
cudaError_t launch_create_nmap(const float * vmap, size_t vp, int rows, int cols, float * nmap, size_t np, cudaStream_t s);
cudaError_t launch_transform_maps(const float * vsrc, const float * nsrc, size_t sp, int rows, int cols, const float * R, const float * t,
                                  float * vdst, float * ndst, size_t dp, cudaStream_t s);
cudaError_t launch_copy_maps(const float * v4, const float * n4, int rows, int cols, float * vdst, float * ndst, size_t dp, cudaStream_t s);
cudaError_t launch_resize_map(const float * in, size_t ip, int srows, int scols, float * out, size_t op, int normalize, cudaStream_t s);
cudaError_t launch_vertices_to_depth(const float * v4, int rows, int cols, float cutoff, float * dst, size_t dp, cudaStream_t s);
cudaError_t launch_z_to_depth(const float * z, int rows, int cols, float cutoff, float * dst, size_t dp, cudaStream_t s);
cudaError_t launch_extract_z(const float * v4, int n, float * z, cudaStream_t s);
cudaError_t launch_pyr_down_gauss_f32(const float * src, size_t sp, int srows, int scols, float * dst, size_t dp, cudaStream_t s);
cudaError_t launch_pyr_down_gauss_u8(const uint8_t * src, size_t sp, int srows, int scols, uint8_t * dst, size_t dp, cudaStream_t s);
cudaError_t launch_bgr_to_intensity(const uint8_t * rgba, size_t sp, int rows, int cols, uint8_t * dst, size_t dp, cudaStream_t s);
cudaError_t launch_derivative_images(const uint8_t * src, size_t sp, int rows, int cols, int16_t * dx, int16_t * dy, size_t dp, cudaStream_t s);
cudaError_t launch_project_points(const float * depth, size_t dp, int rows, int cols, float fx, float fy, float cx, float cy, float * cloud,
                                  size_t cp, cudaStream_t s);

// ---- the step before the tracker (ef_ops_depth.cu): ElasticFusion::filterDepth / metriciseDepth ----
cudaError_t launch_depth_bilateral(const uint16_t * src, size_t sp, int rows, int cols, float max_depth_m, uint16_t * dst, size_t dp, cudaStream_t s);
cudaError_t launch_depth_metric(const uint16_t * src, size_t sp, int rows, int cols, float max_depth_m, float * dst, size_t dp, cudaStream_t s);

// ---- the step around the tracker (ef_ops_predict.cu): IndexMap::combinedPredict + FillIn ----
struct SplatArgs
{
    const float * surfels;
    size_t stride_bytes;
    int count;
    float t_inv[12];
    float cx, cy, fx, fy;
    int rows, cols;
    float max_depth, conf_threshold;
    int time, max_time, time_delta;
    void * keys; // rows * cols * 24 bytes: depth keys, then the per-pixel viewing rays
    uint8_t * image;
    float * vertex, * normal;
    uint16_t * time_out;
    uint8_t * inst = nullptr; // InstanceFusion's fifth target (instance colour, rgba8); may be null
};
cudaError_t launch_splat_predict(const SplatArgs & a, cudaStream_t s);
cudaError_t launch_fill_vertex(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                               float * out, cudaStream_t s);
cudaError_t launch_fill_normal(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                               float * out, cudaStream_t s);
cudaError_t launch_fill_rgb(const uint8_t * predicted, const uint8_t * raw, int rows, int cols, int passthrough, uint8_t * out, cudaStream_t s);

// ---- fused pyramid builders (ef_build_fused.cu); dense outputs ----
// R == nullptr: no transform (initICP maps overload); else v' = R v + t, n' = R n at every level (initICPModel)
cudaError_t launch_build_maps(const float * v4, const float * n4, int rows, int cols, float * const vmaps[3], float * const nmaps[3], float * tmp_z,
                              const float * R, const float * t, cudaStream_t s);
// vertex + normal map of one level (+ dense copy of the level's depth, + pyrDown into next_depth; either may be null)
cudaError_t launch_depth_level(const uint16_t * depth, size_t dpitch_bytes, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff,
                               float * vmap, float * nmap, uint16_t * depth_copy, uint16_t * next_depth, cudaStream_t s);
// intensity (+ float depth from the kept z channel when depth0 != null) at level 0 and their Gaussian pyrDowns to level 1
// z: the kept z channel (z_stride 1) or the z component of the RGBA32F vertex texture itself (z = &v4[2], z_stride 4)
cudaError_t launch_rgbd_level0(const uint8_t * rgba, size_t pitch_bytes, const float * z, int z_stride, float cutoff, int rows, int cols,
                               uint8_t * img0, float * depth0, uint8_t * img1, float * depth1, cudaStream_t s);
cudaError_t launch_rgbd_level1(const uint8_t * img1, const float * depth1, int rows1, int cols1, uint8_t * img2, float * depth2, cudaStream_t s);
cudaError_t launch_derivatives3(const uint8_t * const img[3], int16_t * const dx[3], int16_t * const dy[3], const int rows[3], const int cols[3],
                                cudaStream_t s);

// every pyramid of a frame-to-model frame (initICPModel + initRGBModel + initICP(depth) + initRGB) in one launch
struct FrameBuildArgs
{
    int rows, cols;
    const float * v4, * n4;      // model vertex / normal textures (RGBA32F)
    const uint8_t * model_rgba, * rgba;
    const uint16_t * depth;
    size_t depth_pitch_bytes;    // 0 = dense
    float R[9], t[3];            // pose of the model maps
    float depth_cutoff, rgb_depth_cutoff;
    float fx[3], fy[3], cx[3], cy[3]; // level intrinsics
    float * vmap_g_prev[3], * nmap_g_prev[3], * tmp_z;
    float * vmap_curr[3], * nmap_curr[3];
    uint16_t * depth_pyr[3];
    uint8_t * next_image[3], * last_image[3];
    float * next_depth[3], * last_depth[3];
};
cudaError_t launch_build_frame(const FrameBuildArgs & a, cudaStream_t s);

// ---- Tier-2 association + reduction operators (ef_ops_reduce.cu) ----
// Scratch block layout (device): [0] ticket (u32) | [128] result (32 floats / 2 ints) | [256] partial rows
constexpr size_t kScratchTicketOff = 0;
constexpr size_t kScratchResultOff = 128;
constexpr size_t kScratchPartialOff = 256;
constexpr int kMaxReduceBlocks = 2048;
constexpr size_t kScratchBytes = kScratchPartialOff + (size_t)kMaxReduceBlocks * 32 * sizeof(float);

// EF_OPT_USE_GRAPH: what changes from one Gauss-Newton iteration to the next, read from device memory by the step
// kernels so that the launches of an iteration are fixed nodes of a replayed CUDA graph
struct IterParams
{
    float Rcurr[9], tcurr[3];  // icpStep
    float Rprev_inv[9], tprev[3]; // icpStep; constant over a frame, but the graph outlives the frame
    float krkinv[9], kt[3];    // computeRgbResidual
    int rgb_only;              // rgbStep: sigma = -1 (RGBDOdometry.cpp:472)
    int pad[3];
};

struct IcpArgs
{
    float Rcurr[9], tcurr[3], Rprev_inv[9], tprev[3];
    float fx, fy, cx, cy; // level intrinsics
    float dist_thresh, angle_thresh;
    int rows, cols;
    const float * vmap_curr, * nmap_curr, * vmap_g_prev, * nmap_g_prev;
    size_t pitch; // bytes, same for the four maps (0 = dense)
};
// result: 29 floats at scratch + kScratchResultOff
cudaError_t launch_icp_step(const IcpArgs & a, void * scratch, cudaStream_t s, const IterParams * it = nullptr);

struct RgbResArgs
{
    float min_scale, max_depth_delta;
    float kt[3], krkinv[9];
    int rows, cols;
    const int16_t * dIdx, * dIdy;
    size_t d_pitch;
    const float * last_depth, * next_depth;
    size_t depth_pitch;
    const uint8_t * last_image, * next_image;
    size_t image_pitch;
    void * corres; // rows*cols 16-byte records, linear
};
// result: int2 {count, sigma} at scratch + kScratchResultOff
cudaError_t launch_rgb_residual(const RgbResArgs & a, void * scratch, cudaStream_t s, const IterParams * it = nullptr);

struct RgbStepArgs
{
    const void * corres;
    float sigma;
    const float * cloud;
    size_t cloud_pitch;
    float fx, fy;
    const int16_t * dIdx, * dIdy;
    size_t d_pitch;
    float sobel_scale;
    int rows, cols;
};
// it != null: sigma is derived on the device from `residual` = the {count, sum} computeRgbResidual produced
cudaError_t launch_rgb_step(const RgbStepArgs & a, void * scratch, cudaStream_t s, const IterParams * it = nullptr, const int * residual = nullptr);

// ---- EF_SOLVE_HOST, fused iteration (ef_iter_fused.cu): gates once per call, then two launches per Gauss-Newton iteration ----
struct HmGatesArgs
{
    const int16_t * dIdx[3], * dIdy[3];
    const float * next_depth[3];
    const uint8_t * next_image[3];
    float * gate_depth[3]; // out: nextDepth where the pixel passes the iteration-invariant photometric gates, NaN elsewhere
    float min_scale[3];
    int rows[3], cols[3];
};
cudaError_t launch_hm_gates(const HmGatesArgs & a, cudaStream_t s);
// 8-byte correspondences rec0 / rec1 (one word each per pixel); {count, sum} at scratch + kScratchResultOff.  cols % 4 == 0, dense rows
cudaError_t launch_hm_residual(const RgbResArgs & a, const float * gate_depth, unsigned * rec0, float * rec1, void * scratch, cudaStream_t s);
struct HmStepArgs
{
    IcpArgs icp;           // maps, pose, level intrinsics, thresholds (dense rows)
    float sobel_scale;
    const unsigned * rec0;
    const float * rec1;
    const uint8_t * next_image;
    const int16_t * dIdx, * dIdy;
    const int * residual;  // device {count, sum} of launch_hm_residual
    int do_icp, do_rgb, rgb_only;
    float * h_out;         // mapped pinned, 65 words: [0, 29) ICP sums | [32, 61) RGB sums | [62] count [63] sum | [64] = seq when done
    unsigned seq;
};
cudaError_t launch_hm_step(const HmStepArgs & a, void * scratch, cudaStream_t s);

struct So3Args
{
    const uint8_t * last_image, * next_image;
    size_t image_pitch;
    float image_basis[9], kinv[9], krlr[9];
    int rows, cols;
};
// result: 11 floats
cudaError_t launch_so3_step(const So3Args & a, void * scratch, cudaStream_t s);

} // namespace ef
