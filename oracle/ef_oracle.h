/*
 * oracle/ef_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C + OpenMP) of the dense frame-to-model tracker of
 * ElasticFusion as shipped in Fancomi2017/InstanceFusion:
 *   elasticfusionpublic/Core/src/Cuda/cudafuncs.cu   (image / pyramid kernels)
 *   elasticfusionpublic/Core/src/Cuda/reduce.cu      (association + reduction kernels)
 *   elasticfusionpublic/Core/src/Utils/RGBDOdometry.cpp, OdometryProvider.h (host loop)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (instancefusion_b200/libef_track.so) never does.
 *
 * Parity status: the reference ships NO golden vectors for this path, so
 * this restatement is pinned against outputs of the reference's own CUDA kernels built
 * unmodified into oracle/_ref/libef_ref.so and run on a B200 (tests/golden/, generated
 * by tests/golden/make_golden.py).  Known, documented deviations of a CPU restatement:
 * the GPU uses approximate division / rsqrt / sqrt (--prec-div=false --prec-sqrt=false),
 * this file uses IEEE operations, so float outputs agree to a few ulp and integer-valued
 * outputs may differ by one LSB where the quotient sits on an integer boundary.
 *
 * Layout conventions (all dense, row-major):
 *   "map3"  = 3-plane SoA float image, 3*rows x cols, component c of pixel (x,y) at row
 *             y + c*rows  (RGBDOdometry.cpp:97-101)
 *   "rgba32f" = interleaved 4 floats per pixel (GL RGBA32F texture contents)
 */
#ifndef EF_ORACLE_H_
#define EF_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudafuncs.cu:57-107 */
void efo_pyr_down_u16(const uint16_t * src, int srows, int scols, uint16_t * dst);
/* cudafuncs.cu:109-149 ; intr already scaled to the level */
void efo_create_vmap(const uint16_t * depth, int rows, int cols, float fx, float fy, float cx, float cy,
                     float cutoff, float * vmap3);
/* cudafuncs.cu:151-204 */
void efo_create_nmap(const float * vmap3, int rows, int cols, float * nmap3);
/* cudafuncs.cu:206-268 ; R row-major 3x3; src may alias dst */
void efo_transform_maps(const float * vsrc3, const float * nsrc3, int rows, int cols, const float * R, const float * t,
                        float * vdst3, float * ndst3);
/* cudafuncs.cu:270-330 */
void efo_copy_maps(const float * v_rgba32f, const float * n_rgba32f, int rows, int cols, float * vmap3, float * nmap3);
/* cudafuncs.cu:365-444 ; in: srows x scols, out: srows/2 x scols/2 */
void efo_resize_map(const float * in3, int srows, int scols, float * out3, int normalize);
/* cudafuncs.cu:526-546 */
void efo_vertices_to_depth(const float * v_rgba32f, int rows, int cols, float cutoff, float * dst);
/* cudafuncs.cu:332-363, 446-468 */
void efo_pyr_down_gauss_f32(const float * src, int srows, int scols, float * dst);
/* cudafuncs.cu:470-524 */
void efo_pyr_down_gauss_u8(const uint8_t * src, int srows, int scols, uint8_t * dst);
/* cudafuncs.cu:548-577 */
void efo_bgr_to_intensity(const uint8_t * rgba8, int rows, int cols, uint8_t * dst);
/* cudafuncs.cu:580-639 */
void efo_derivative_images(const uint8_t * src, int rows, int cols, int16_t * dx, int16_t * dy);
/* cudafuncs.cu:641-674 ; intr already scaled to the level; cloud = rows x cols x 3 floats */
/* the step before the tracker (GLSL in the reference): Shaders/depth_bilateral.frag:30-76, depth_metric.frag:28-40 */
void efo_depth_bilateral(const uint16_t * src, int rows, int cols, float max_depth_m, uint16_t * dst);
void efo_depth_metric(const uint16_t * src, int rows, int cols, float max_depth_m, float * dst);

/* the step around the tracker (OpenGL in the reference): IndexMap::combinedPredict (Shaders/splat.vert, combo_splat.frag)
 * and FillIn (Shaders/fill_vertex.frag, fill_normal.frag, fill_rgb.frag).  Parity against GL is unpinned; see ef_oracle.c */
void efo_splat_predict(const float * surfels, int stride_floats, int count, const float * t_inv16, float cx, float cy, float fx, float fy, int rows,
                       int cols, float max_depth, float conf_threshold, int time, int max_time, int time_delta, uint8_t * image_rgba8,
                       float * vertex_rgba32f, float * normal_rgba32f, uint16_t * time_u16);
void efo_fill_vertex(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                     float * out);
void efo_fill_normal(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                     float * out);
void efo_fill_rgb(const uint8_t * predicted, const uint8_t * raw, int rows, int cols, int passthrough, uint8_t * out);

void efo_project_point_cloud(const float * depth, int rows, int cols, float fx, float fy, float cx, float cy, float * cloud);

/* reduce.cu:257-490.  out29 = 27 upper-triangle products of [J|r] + residual + inliers,
 * accumulated in double from per-pixel float products, rounded to float at the end. */
void efo_icp_step(const float * Rcurr, const float * tcurr, const float * vmap_curr3, const float * nmap_curr3,
                  const float * Rprev_inv, const float * tprev, float fx, float fy, float cx, float cy,
                  const float * vmap_g_prev3, const float * nmap_g_prev3, float dist_thresh, float angle_thresh,
                  int rows, int cols, float * out29);

/* one correspondence record: types.cuh:75-81 */
typedef struct
{
    int16_t zero_x, zero_y;
    int16_t one_x, one_y;
    float diff;
    uint8_t valid;
    uint8_t pad[3];
} efo_data_term;

/* reduce.cu:739-936.  krkinv row-major 3x3, kt[3]; corres = rows*cols records */
void efo_rgb_residual(float min_scale, const int16_t * dIdx, const int16_t * dIdy, const float * last_depth,
                      const float * next_depth, const uint8_t * last_image, const uint8_t * next_image,
                      efo_data_term * corres, float max_depth_delta, const float * kt, const float * krkinv, int rows,
                      int cols, int * sigma_sum, int * count);
/* reduce.cu:494-678 */
void efo_rgb_step(const efo_data_term * corres, float sigma, const float * cloud, float fx, float fy,
                  const int16_t * dIdx, const int16_t * dIdy, float sobel_scale, int rows, int cols, float * out29);
/* reduce.cu:938-1141.  out11 = 9 products + residual + inliers */
void efo_so3_step(const uint8_t * last_image, const uint8_t * next_image, const float * image_basis, const float * kinv,
                  const float * krlr, int rows, int cols, float * out11);

/* unpack 29 (11) floats into row-major A (6x6 / 3x3), b, residual[2]: reduce.cu:475-489, 1126-1140 */
void efo_unpack_se3(const float * out29, float * A36, float * b6, float * residual2);
void efo_unpack_so3(const float * out11, float * A9, float * b3, float * residual2);

/* ---- full tracker: RGBDOdometry.cpp:21-608 ---- */
typedef struct efo_tracker efo_tracker;

typedef struct
{
    float last_icp_error, last_icp_count;
    float last_rgb_error, last_rgb_count;
    float last_so3_error, last_so3_count;
    double last_A[36];
    double last_b[6];
    int so3_iterations;
    int se3_iterations[3];
} efo_stats;

efo_tracker * efo_tracker_create(int width, int height, float cx, float cy, float fx, float fy, float dist_thresh,
                                 float angle_thresh);
void efo_tracker_destroy(efo_tracker * t);
void efo_init_icp_depth(efo_tracker * t, const uint16_t * depth, float cutoff);                        /* :118-142 */
void efo_init_icp_maps(efo_tracker * t, const float * v_rgba32f, const float * n_rgba32f, float cutoff); /* :144-167 */
void efo_init_icp_model(efo_tracker * t, const float * v_rgba32f, const float * n_rgba32f, float cutoff,
                        const float * pose16);                                                          /* :169-206 */
void efo_init_rgb(efo_tracker * t, const uint8_t * rgba8);                                              /* :243-247 */
void efo_init_rgb_model(efo_tracker * t, const uint8_t * rgba8);                                        /* :237-241 */
void efo_init_first_rgb(efo_tracker * t, const uint8_t * rgba8);                                        /* :249-265 */
void efo_get_incremental_transformation(efo_tracker * t, float * trans3, float * rot9, int rgb_only, float icp_weight,
                                        int pyramid, int fast_odom, int so3, efo_stats * stats);      /* :267-603 */
void efo_get_covariance(const efo_tracker * t, double * cov36);                                         /* :605-608 */

/* access to internal pyramids for tests; name in {"vmap_curr","nmap_curr","vmap_g_prev","nmap_g_prev",
 * "last_depth","next_depth","last_image","next_image","last_next_image","dIdx","dIdy","depth_tmp"} */
const void * efo_tracker_buffer(const efo_tracker * t, const char * name, int level);

/* host helpers (OdometryProvider.h:35-93 and the Eigen calls in RGBDOdometry.cpp) exposed for tests */
void efo_rodrigues(const double * v3, double * R9);
int efo_ldlt_solve_f64(const double * A, const double * b, int n, double * x);
void efo_inverse4_f64(const double * M16, double * Minv16);
void efo_ldlt_solve3_f32(const float * A9, const float * b3, float * x3);
void efo_mul33_f64(const double * A, const double * B, double * C);
void efo_inverse3_f64(const double * m, double * o);
void efo_inverse3_f32(const float * m, float * o);

#ifdef __cplusplus
}
#endif

#endif /* EF_ORACLE_H_ */
