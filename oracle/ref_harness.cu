/*
 * oracle/ref_harness.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * C-ABI harness around the UNMODIFIED reference CUDA operators
 * (elasticfusionpublic/Core/src/Cuda/cudafuncs.cuh:64-177), linked from the reference's
 * own sources into oracle/_ref/libef_ref.so by oracle/Makefile.  Two tiers:
 *
 *   efr_<op>(...)         one reference operator on host arrays (upload, run, download)
 *   efr_tracker_*(...)    an Eigen-free / GL-free restatement of the host driver
 *                         RGBDOdometry.cpp:21-608 that calls ONLY reference operators, with
 *                         GPUConfig's default launch shapes (GPUConfig.h:53-60).  Its host
 *                         linear algebra comes from oracle/ef_oracle.c (efo_* helpers).
 *
 * This is the parity oracle for the GPU tests and the "reference CUDA on the same B200"
 * timing arm of bench.py.  It needs a GPU; nothing in the product links it.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>

#include "cudafuncs.cuh" /* reference header, found via -I at build time */
#include "ef_oracle.h"

#define EFR_API extern "C" __attribute__((visibility("default")))

namespace
{

template<class T>
void up2d(DeviceArray2D<T> & d, const T * host, int rows, int cols)
{
    d.create(rows, cols);
    cudaMemset2D(d.ptr(), d.step(), 0, cols * sizeof(T), rows);
    if(host) d.upload(host, cols * sizeof(T), rows, cols);
}

template<class T>
void alloc2d(DeviceArray2D<T> & d, int rows, int cols)
{
    d.create(rows, cols);
    cudaMemset2D(d.ptr(), d.step(), 0, cols * sizeof(T), rows);
}

mat33 to_mat33(const float * m)
{
    mat33 r;
    memcpy(&r.data[0], m, sizeof(mat33));
    return r;
}

} // namespace

/* ---------------------------------------------------------------------------------------- */
/* per-operator entry points                                                                */
/* ---------------------------------------------------------------------------------------- */
EFR_API int efr_device_count()
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

EFR_API void efr_pyr_down_u16(const unsigned short * src, int srows, int scols, unsigned short * dst)
{
    DeviceArray2D<unsigned short> s, d;
    up2d(s, src, srows, scols);
    alloc2d(d, srows / 2, scols / 2);
    pyrDown(s, d);
    cudaDeviceSynchronize();
    d.download(dst, (scols / 2) * sizeof(unsigned short));
}

/* vmap is in/out: invalid pixels leave planes y,z untouched */
EFR_API void efr_create_vmap(const unsigned short * depth, int rows, int cols, float fx, float fy, float cx, float cy,
                             float cutoff, float * vmap)
{
    DeviceArray2D<unsigned short> d;
    DeviceArray2D<float> v;
    up2d(d, depth, rows, cols);
    up2d(v, vmap, rows * 3, cols);
    createVMap(CameraModel(fx, fy, cx, cy), d, v, cutoff);
    cudaDeviceSynchronize();
    v.download(vmap, cols * sizeof(float));
}

EFR_API void efr_create_nmap(const float * vmap, int rows, int cols, float * nmap)
{
    DeviceArray2D<float> v, n;
    up2d(v, vmap, rows * 3, cols);
    up2d(n, nmap, rows * 3, cols);
    createNMap(v, n);
    cudaDeviceSynchronize();
    n.download(nmap, cols * sizeof(float));
}

/* in place, like RGBDOdometry.cpp:202 */
EFR_API void efr_transform_maps(float * vmap, float * nmap, int rows, int cols, const float * R, const float * t)
{
    DeviceArray2D<float> v, n;
    up2d(v, vmap, rows * 3, cols);
    up2d(n, nmap, rows * 3, cols);
    float3 tv = {t[0], t[1], t[2]};
    tranformMaps(v, n, to_mat33(R), tv, v, n);
    cudaDeviceSynchronize();
    v.download(vmap, cols * sizeof(float));
    n.download(nmap, cols * sizeof(float));
}

EFR_API void efr_copy_maps(const float * v4, const float * n4, int rows, int cols, float * vmap, float * nmap)
{
    DeviceArray<float> vs, ns;
    vs.upload(v4, (size_t)rows * cols * 4);
    ns.upload(n4, (size_t)rows * cols * 4);
    DeviceArray2D<float> v, n;
    alloc2d(v, rows * 3, cols);
    alloc2d(n, rows * 3, cols);
    copyMaps(vs, ns, v, n);
    cudaDeviceSynchronize();
    v.download(vmap, cols * sizeof(float));
    n.download(nmap, cols * sizeof(float));
}

/* out is in/out (invalid pixels leave y,z untouched) */
EFR_API void efr_resize_map(const float * in, int srows, int scols, float * out, int normalize)
{
    DeviceArray2D<float> i, o;
    up2d(i, in, srows * 3, scols);
    up2d(o, out, (srows / 2) * 3, scols / 2);
    if(normalize) resizeNMap(i, o);
    else resizeVMap(i, o);
    cudaDeviceSynchronize();
    o.download(out, (scols / 2) * sizeof(float));
}

EFR_API void efr_vertices_to_depth(const float * v4, int rows, int cols, float cutoff, float * dst)
{
    DeviceArray<float> vs;
    vs.upload(v4, (size_t)rows * cols * 4);
    DeviceArray2D<float> d;
    alloc2d(d, rows, cols);
    verticesToDepth(vs, d, cutoff);
    cudaDeviceSynchronize();
    d.download(dst, cols * sizeof(float));
}

EFR_API void efr_pyr_down_gauss_f32(const float * src, int srows, int scols, float * dst)
{
    DeviceArray2D<float> s, d;
    up2d(s, src, srows, scols);
    alloc2d(d, srows / 2, scols / 2);
    pyrDownGaussF(s, d);
    cudaDeviceSynchronize();
    d.download(dst, (scols / 2) * sizeof(float));
}

EFR_API void efr_pyr_down_gauss_u8(const unsigned char * src, int srows, int scols, unsigned char * dst)
{
    DeviceArray2D<unsigned char> s, d;
    up2d(s, src, srows, scols);
    alloc2d(d, srows / 2, scols / 2);
    pyrDownUcharGauss(s, d);
    cudaDeviceSynchronize();
    d.download(dst, (scols / 2));
}

static cudaArray * make_rgba_array(const unsigned char * rgba, int rows, int cols, bool from_device)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    cudaArray * arr = NULL;
    cudaMallocArray(&arr, &desc, cols, rows);
    cudaMemcpy2DToArray(arr, 0, 0, rgba, cols * 4, cols * 4, rows, from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
    return arr;
}

EFR_API void efr_bgr_to_intensity(const unsigned char * rgba, int rows, int cols, unsigned char * dst)
{
    cudaArray * arr = make_rgba_array(rgba, rows, cols, false);
    DeviceArray2D<unsigned char> d;
    alloc2d(d, rows, cols);
    imageBGRToIntensity(arr, d);
    cudaDeviceSynchronize();
    d.download(dst, cols);
    cudaFreeArray(arr);
}

EFR_API void efr_derivative_images(const unsigned char * src, int rows, int cols, short * dx, short * dy)
{
    DeviceArray2D<unsigned char> s;
    DeviceArray2D<short> ddx, ddy;
    up2d(s, src, rows, cols);
    alloc2d(ddx, rows, cols);
    alloc2d(ddy, rows, cols);
    computeDerivativeImages(s, ddx, ddy);
    ddx.download(dx, cols * sizeof(short));
    ddy.download(dy, cols * sizeof(short));
}

/* intr = level-0 intrinsics, level applied inside as in the reference */
EFR_API void efr_project_point_cloud(const float * depth, int rows, int cols, float fx, float fy, float cx, float cy, int level,
                                     float * cloud)
{
    DeviceArray2D<float> d;
    DeviceArray2D<float3> c;
    up2d(d, depth, rows, cols);
    alloc2d(c, rows, cols);
    CameraModel intr(fx, fy, cx, cy);
    projectToPointCloud(d, c, intr, level);
    c.download(cloud, cols * sizeof(float3));
}

EFR_API void efr_icp_step(const float * Rcurr, const float * tcurr, const float * vmap_curr, const float * nmap_curr,
                          const float * Rprev_inv, const float * tprev, float fx, float fy, float cx, float cy,
                          const float * vmap_g_prev, const float * nmap_g_prev, float dist_thresh, float angle_thresh, int rows,
                          int cols, int threads, int blocks, float * A36, float * b6, float * residual2)
{
    DeviceArray2D<float> vc, nc, vp, np;
    up2d(vc, vmap_curr, rows * 3, cols);
    up2d(nc, nmap_curr, rows * 3, cols);
    up2d(vp, vmap_g_prev, rows * 3, cols);
    up2d(np, nmap_g_prev, rows * 3, cols);
    DeviceArray<JtJJtrSE3> sum, out;
    sum.create(MAX_THREADS);
    out.create(1);
    float3 tc = {tcurr[0], tcurr[1], tcurr[2]}, tp = {tprev[0], tprev[1], tprev[2]};
    icpStep(to_mat33(Rcurr), tc, vc, nc, to_mat33(Rprev_inv), tp, CameraModel(fx, fy, cx, cy), vp, np, dist_thresh,
            angle_thresh, sum, out, A36, b6, residual2, threads, blocks);
}

/* corres: rows*cols records of 16 bytes (types.cuh:75-81) */
EFR_API void efr_rgb_residual(float min_scale, const short * dIdx, const short * dIdy, const float * last_depth,
                              const float * next_depth, const unsigned char * last_image, const unsigned char * next_image,
                              void * corres, float max_depth_delta, const float * kt, const float * krkinv, int rows, int cols,
                              int threads, int blocks, int * sigma_sum, int * count)
{
    DeviceArray2D<short> dx, dy;
    DeviceArray2D<float> ld, nd;
    DeviceArray2D<unsigned char> li, ni;
    DeviceArray2D<DataTerm> cor;
    up2d(dx, dIdx, rows, cols);
    up2d(dy, dIdy, rows, cols);
    up2d(ld, last_depth, rows, cols);
    up2d(nd, next_depth, rows, cols);
    up2d(li, last_image, rows, cols);
    up2d(ni, next_image, rows, cols);
    alloc2d(cor, rows, cols);
    DeviceArray<int2> sum;
    sum.create(MAX_THREADS);
    float3 ktv = {kt[0], kt[1], kt[2]};
    computeRgbResidual(min_scale, dx, dy, ld, nd, li, ni, cor, sum, max_depth_delta, ktv, to_mat33(krkinv), *sigma_sum, *count,
                       threads, blocks);
    if(corres) cor.download(corres, cols * sizeof(DataTerm));
}

EFR_API void efr_rgb_step(const void * corres, float sigma, const float * cloud, float fx, float fy, const short * dIdx,
                          const short * dIdy, float sobel_scale, int rows, int cols, int threads, int blocks, float * A36,
                          float * b6)
{
    DeviceArray2D<DataTerm> cor;
    DeviceArray2D<float3> cl;
    DeviceArray2D<short> dx, dy;
    up2d(cor, (const DataTerm *)corres, rows, cols);
    up2d(cl, (const float3 *)cloud, rows, cols);
    up2d(dx, dIdx, rows, cols);
    up2d(dy, dIdy, rows, cols);
    DeviceArray<JtJJtrSE3> sum, out;
    sum.create(MAX_THREADS);
    out.create(1);
    rgbStep(cor, sigma, cl, fx, fy, dx, dy, sobel_scale, sum, out, A36, b6, threads, blocks);
}

EFR_API void efr_so3_step(const unsigned char * last_image, const unsigned char * next_image, const float * image_basis,
                          const float * kinv, const float * krlr, int rows, int cols, int threads, int blocks, float * A9,
                          float * b3, float * residual2)
{
    DeviceArray2D<unsigned char> li, ni;
    up2d(li, last_image, rows, cols);
    up2d(ni, next_image, rows, cols);
    DeviceArray<JtJJtrSO3> sum, out;
    sum.create(MAX_THREADS);
    out.create(1);
    so3Step(li, ni, to_mat33(image_basis), to_mat33(kinv), to_mat33(krlr), sum, out, A9, b3, residual2, threads, blocks);
}

/* ---------------------------------------------------------------------------------------- */
/* tracker: RGBDOdometry.cpp restated over the reference operators                          */
/* ---------------------------------------------------------------------------------------- */
static const int NUM_PYRS = 3;

struct efr_tracker
{
    int width, height;
    CameraModel intr;
    float dist_thresh, angle_thresh;
    float sobel_scale, max_depth_delta_rgb, max_depth_rgb;
    float min_grad[NUM_PYRS];

    /* GPUConfig.h:53-60 defaults (the lookup by device name has no "NVIDIA B200" entry) */
    int icp_threads, icp_blocks, rgb_threads, rgb_blocks, rgbres_threads, rgbres_blocks, so3_threads, so3_blocks;

    std::vector<DeviceArray2D<unsigned short> > depth_tmp;
    DeviceArray<float> vmaps_tmp, nmaps_tmp;
    std::vector<DeviceArray2D<float> > vmaps_g_prev, nmaps_g_prev, vmaps_curr, nmaps_curr;
    DeviceArray<JtJJtrSE3> sumDataSE3, outDataSE3;
    DeviceArray<int2> sumResidualRGB;
    DeviceArray<JtJJtrSO3> sumDataSO3, outDataSO3;
    DeviceArray2D<float> lastDepth[NUM_PYRS], nextDepth[NUM_PYRS];
    DeviceArray2D<unsigned char> lastImage[NUM_PYRS], nextImage[NUM_PYRS], lastNextImage[NUM_PYRS];
    DeviceArray2D<short> nextdIdx[NUM_PYRS], nextdIdy[NUM_PYRS];
    DeviceArray2D<DataTerm> corresImg[NUM_PYRS];
    DeviceArray2D<float3> pointClouds[NUM_PYRS];
    cudaArray * rgb_array; /* stands in for the GL texture's cudaArray */

    efo_stats st;
};

EFR_API efr_tracker * efr_tracker_create(int width, int height, float cx, float cy, float fx, float fy, float dist_thresh,
                                         float angle_thresh)
{
    efr_tracker * t = new efr_tracker();
    t->width = width; t->height = height;
    t->intr = CameraModel(fx, fy, cx, cy);
    t->dist_thresh = dist_thresh; t->angle_thresh = angle_thresh;
    t->sobel_scale = (float)(1.0 / pow(2.0, 3));
    t->max_depth_delta_rgb = 0.07f;
    t->max_depth_rgb = 6.0f;
    t->min_grad[0] = 5; t->min_grad[1] = 3; t->min_grad[2] = 1;
    t->icp_threads = 128; t->icp_blocks = 112;
    t->rgb_threads = 128; t->rgb_blocks = 112;
    t->rgbres_threads = 256; t->rgbres_blocks = 336;
    t->so3_threads = 160; t->so3_blocks = 64;
    memset(&t->st, 0, sizeof(t->st));
    t->st.last_icp_count = t->st.last_rgb_count = t->st.last_so3_count = (float)(width * height);

    t->sumDataSE3.create(MAX_THREADS);
    t->outDataSE3.create(1);
    t->sumResidualRGB.create(MAX_THREADS);
    t->sumDataSO3.create(MAX_THREADS);
    t->outDataSO3.create(1);
    t->depth_tmp.resize(NUM_PYRS);
    t->vmaps_g_prev.resize(NUM_PYRS); t->nmaps_g_prev.resize(NUM_PYRS);
    t->vmaps_curr.resize(NUM_PYRS); t->nmaps_curr.resize(NUM_PYRS);
    for(int i = 0; i < NUM_PYRS; i++)
    {
        int r = height >> i, c = width >> i;
        alloc2d(t->lastDepth[i], r, c); alloc2d(t->lastImage[i], r, c);
        alloc2d(t->nextDepth[i], r, c); alloc2d(t->nextImage[i], r, c);
        alloc2d(t->lastNextImage[i], r, c);
        alloc2d(t->nextdIdx[i], r, c); alloc2d(t->nextdIdy[i], r, c);
        alloc2d(t->pointClouds[i], r, c); alloc2d(t->corresImg[i], r, c);
        alloc2d(t->depth_tmp[i], r, c);
        alloc2d(t->vmaps_g_prev[i], r * 3, c); alloc2d(t->nmaps_g_prev[i], r * 3, c);
        alloc2d(t->vmaps_curr[i], r * 3, c); alloc2d(t->nmaps_curr[i], r * 3, c);
    }
    t->vmaps_tmp.create((size_t)height * 4 * width);
    t->nmaps_tmp.create((size_t)height * 4 * width);
    cudaMemset(t->vmaps_tmp.ptr(), 0, t->vmaps_tmp.sizeBytes());
    cudaMemset(t->nmaps_tmp.ptr(), 0, t->nmaps_tmp.sizeBytes());
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    cudaMallocArray(&t->rgb_array, &desc, width, height);
    cudaDeviceSynchronize();
    return t;
}

EFR_API void efr_tracker_destroy(efr_tracker * t)
{
    if(!t) return;
    cudaFreeArray(t->rgb_array);
    delete t;
}

EFR_API void efr_tracker_set_config(efr_tracker * t, int icp_threads, int icp_blocks, int rgb_threads, int rgb_blocks,
                                    int rgbres_threads, int rgbres_blocks, int so3_threads, int so3_blocks)
{
    t->icp_threads = icp_threads; t->icp_blocks = icp_blocks;
    t->rgb_threads = rgb_threads; t->rgb_blocks = rgb_blocks;
    t->rgbres_threads = rgbres_threads; t->rgbres_blocks = rgbres_blocks;
    t->so3_threads = so3_threads; t->so3_blocks = so3_blocks;
}

static cudaMemcpyKind kind_of(int on_device) { return on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }

/* RGBDOdometry.cpp:118-142 ; the GL texture copy becomes a 2-D copy from a linear buffer */
EFR_API void efr_init_icp_depth(efr_tracker * t, const unsigned short * depth, float cutoff, int on_device)
{
    DeviceArray2D<unsigned short> & d0 = t->depth_tmp[0];
    cudaMemcpy2D(d0.ptr(0), d0.step(), depth, t->width * sizeof(unsigned short), d0.colsBytes(), d0.rows(), kind_of(on_device));
    for(int i = 1; i < NUM_PYRS; ++i) pyrDown(t->depth_tmp[i - 1], t->depth_tmp[i]);
    for(int i = 0; i < NUM_PYRS; ++i)
    {
        createVMap(t->intr(i), t->depth_tmp[i], t->vmaps_curr[i], cutoff);
        createNMap(t->vmaps_curr[i], t->nmaps_curr[i]);
    }
    cudaDeviceSynchronize();
}

/* RGBDOdometry.cpp:144-167 */
EFR_API void efr_init_icp_maps(efr_tracker * t, const float * v4, const float * n4, float cutoff, int on_device)
{
    (void)cutoff;
    cudaMemcpy(t->vmaps_tmp.ptr(), v4, t->vmaps_tmp.sizeBytes(), kind_of(on_device));
    cudaMemcpy(t->nmaps_tmp.ptr(), n4, t->nmaps_tmp.sizeBytes(), kind_of(on_device));
    copyMaps(t->vmaps_tmp, t->nmaps_tmp, t->vmaps_curr[0], t->nmaps_curr[0]);
    for(int i = 1; i < NUM_PYRS; ++i)
    {
        resizeVMap(t->vmaps_curr[i - 1], t->vmaps_curr[i]);
        resizeNMap(t->nmaps_curr[i - 1], t->nmaps_curr[i]);
    }
    cudaDeviceSynchronize();
}

/* RGBDOdometry.cpp:169-206 ; pose row-major 4x4 */
EFR_API void efr_init_icp_model(efr_tracker * t, const float * v4, const float * n4, float cutoff, const float * pose,
                                int on_device)
{
    (void)cutoff;
    cudaMemcpy(t->vmaps_tmp.ptr(), v4, t->vmaps_tmp.sizeBytes(), kind_of(on_device));
    cudaMemcpy(t->nmaps_tmp.ptr(), n4, t->nmaps_tmp.sizeBytes(), kind_of(on_device));
    copyMaps(t->vmaps_tmp, t->nmaps_tmp, t->vmaps_g_prev[0], t->nmaps_g_prev[0]);
    for(int i = 1; i < NUM_PYRS; ++i)
    {
        resizeVMap(t->vmaps_g_prev[i - 1], t->vmaps_g_prev[i]);
        resizeNMap(t->nmaps_g_prev[i - 1], t->nmaps_g_prev[i]);
    }
    float R[9] = {pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10]};
    float3 tv = {pose[3], pose[7], pose[11]};
    mat33 dR = to_mat33(R);
    for(int i = 0; i < NUM_PYRS; ++i)
        tranformMaps(t->vmaps_g_prev[i], t->nmaps_g_prev[i], dR, tv, t->vmaps_g_prev[i], t->nmaps_g_prev[i]);
    cudaDeviceSynchronize();
}

/* RGBDOdometry.cpp:208-235 */
static void populate(efr_tracker * t, const unsigned char * rgba, int on_device, DeviceArray2D<float> * depths,
                     DeviceArray2D<unsigned char> * images)
{
    verticesToDepth(t->vmaps_tmp, depths[0], t->max_depth_rgb);
    for(int i = 0; i + 1 < NUM_PYRS; i++) pyrDownGaussF(depths[i], depths[i + 1]);
    cudaMemcpy2DToArray(t->rgb_array, 0, 0, rgba, t->width * 4, t->width * 4, t->height, kind_of(on_device));
    imageBGRToIntensity(t->rgb_array, images[0]);
    for(int i = 0; i + 1 < NUM_PYRS; i++) pyrDownUcharGauss(images[i], images[i + 1]);
    cudaDeviceSynchronize();
}

EFR_API void efr_init_rgb_model(efr_tracker * t, const unsigned char * rgba, int on_device)
{
    populate(t, rgba, on_device, &t->lastDepth[0], &t->lastImage[0]);
}

EFR_API void efr_init_rgb(efr_tracker * t, const unsigned char * rgba, int on_device)
{
    populate(t, rgba, on_device, &t->nextDepth[0], &t->nextImage[0]);
}

/* RGBDOdometry.cpp:249-265 */
EFR_API void efr_init_first_rgb(efr_tracker * t, const unsigned char * rgba, int on_device)
{
    cudaMemcpy2DToArray(t->rgb_array, 0, 0, rgba, t->width * 4, t->width * 4, t->height, kind_of(on_device));
    imageBGRToIntensity(t->rgb_array, t->lastNextImage[0]);
    for(int i = 0; i + 1 < NUM_PYRS; i++) pyrDownUcharGauss(t->lastNextImage[i], t->lastNextImage[i + 1]);
}

static void cast9(const double * d, float * f) { for(int i = 0; i < 9; i++) f[i] = (float)d[i]; }

/* RGBDOdometry.cpp:267-603 */
EFR_API void efr_get_incremental_transformation(efr_tracker * t, float * trans, float * rot, int rgb_only, float icp_weight,
                                                int pyramid, int fast_odom, int so3, efo_stats * stats)
{
    const bool icp = !rgb_only && icp_weight > 0;
    const bool rgb = rgb_only || icp_weight < 100;

    float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
    memcpy(Rprev, rot, sizeof(Rprev)); memcpy(tprev, trans, sizeof(tprev));
    memcpy(Rcurr, rot, sizeof(Rcurr)); memcpy(tcurr, trans, sizeof(tcurr));

    if(rgb)
        for(int i = 0; i < NUM_PYRS; i++) computeDerivativeImages(t->nextImage[i], t->nextdIdx[i], t->nextdIdy[i]);

    double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    t->st.so3_iterations = 0;
    t->st.se3_iterations[0] = t->st.se3_iterations[1] = t->st.se3_iterations[2] = 0;

    if(so3)
    {
        const int lvl = 2;
        float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        CameraModel li = t->intr(lvl);
        double K[9] = {li.fx, 0, li.cx, 0, li.fy, li.cy, 0, 0, 1}, K_inv[9];
        efo_inverse3_f64(K, K_inv);
        float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
        double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

        for(int i = 0; i < 10; i++)
        {
            double tmp[9], Hd[9];
            efo_mul33_f64(K, resultR, tmp);
            efo_mul33_f64(tmp, K_inv, Hd);
            float H[9], kinv[9], krlr[9];
            cast9(Hd, H); cast9(K_inv, kinv); cast9(tmp, krlr);

            float jtj[9], jtr[3], residual[2];
            so3Step(t->lastNextImage[lvl], t->nextImage[lvl], to_mat33(H), to_mat33(kinv), to_mat33(krlr), t->sumDataSO3,
                    t->outDataSO3, jtj, jtr, residual, t->so3_threads, t->so3_blocks);
            t->st.so3_iterations++;

            t->st.last_so3_error = sqrtf(residual[0]) / residual[1];
            t->st.last_so3_count = residual[1];
            if(t->st.last_so3_error < lastError && lastCount == t->st.last_so3_count) break;
            else if(t->st.last_so3_error > lastError + 0.001)
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
            efo_ldlt_solve3_f32(jtj, jtr, delta);
            double dd[3] = {delta[0], delta[1], delta[2]}, rotUpdate[9];
            efo_rodrigues(dd, rotUpdate);
            float ru[9], nr[9];
            cast9(rotUpdate, ru);
            for(int r = 0; r < 3; r++)
                for(int c = 0; c < 3; c++)
                    nr[r * 3 + c] = ru[r * 3] * R_lr[c] + ru[r * 3 + 1] * R_lr[3 + c] + ru[r * 3 + 2] * R_lr[6 + c];
            memcpy(R_lr, nr, sizeof(nr));
            for(int k = 0; k < 9; k++) resultR[k] = R_lr[k];
        }
    }

    int iterations[NUM_PYRS] = {fast_odom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0};

    float Rprev_inv[9];
    efo_inverse3_f32(Rprev, Rprev_inv);
    mat33 device_Rprev_inv = to_mat33(Rprev_inv);
    float3 device_tprev = {tprev[0], tprev[1], tprev[2]};

    double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if(so3)
        for(int x = 0; x < 3; x++)
            for(int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

    for(int i = NUM_PYRS - 1; i >= 0; i--)
    {
        if(rgb) projectToPointCloud(t->lastDepth[i], t->pointClouds[i], t->intr, i);

        CameraModel li = t->intr(i);
        double K[9] = {li.fx, 0, li.cx, 0, li.fy, li.cy, 0, 0, 1}, K_inv[9];
        efo_inverse3_f64(K, K_inv);
        t->st.last_rgb_error = FLT_MAX;

        for(int j = 0; j < iterations[i]; j++)
        {
            double Rt[16];
            efo_inverse4_f64(resultRt, Rt);
            double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
            double tmp[9], KRK_inv[9];
            efo_mul33_f64(K, R, tmp);
            efo_mul33_f64(tmp, K_inv, KRK_inv);
            float krkInv[9];
            cast9(KRK_inv, krkInv);
            double tv[3] = {Rt[3], Rt[7], Rt[11]};
            float3 kt;
            kt.x = (float)(K[0] * tv[0] + K[1] * tv[1] + K[2] * tv[2]);
            kt.y = (float)(K[3] * tv[0] + K[4] * tv[1] + K[5] * tv[2]);
            kt.z = (float)(K[6] * tv[0] + K[7] * tv[1] + K[8] * tv[2]);

            int sigma = 0, rgbSize = 0;
            if(rgb)
                computeRgbResidual((float)(pow(t->min_grad[i], 2.0) / pow(t->sobel_scale, 2.0)), t->nextdIdx[i], t->nextdIdy[i],
                                   t->lastDepth[i], t->nextDepth[i], t->lastImage[i], t->nextImage[i], t->corresImg[i],
                                   t->sumResidualRGB, t->max_depth_delta_rgb, kt, to_mat33(krkInv), sigma, rgbSize,
                                   t->rgbres_threads, t->rgbres_blocks);

            float sigmaVal = std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize);
            float rgbError = std::sqrt(sigma) / (rgbSize == 0 ? 1 : rgbSize);
            if(rgb_only && rgbError > t->st.last_rgb_error) break;
            t->st.last_rgb_error = rgbError;
            t->st.last_rgb_count = rgbSize;
            if(rgb_only) sigmaVal = -1;

            float A_icp[36] = {0}, b_icp[6] = {0}, A_rgb[36] = {0}, b_rgb[6] = {0}, residual[2];
            float3 device_tcurr = {tcurr[0], tcurr[1], tcurr[2]};
            if(icp)
            {
                icpStep(to_mat33(Rcurr), device_tcurr, t->vmaps_curr[i], t->nmaps_curr[i], device_Rprev_inv, device_tprev,
                        t->intr(i), t->vmaps_g_prev[i], t->nmaps_g_prev[i], t->dist_thresh, t->angle_thresh, t->sumDataSE3,
                        t->outDataSE3, A_icp, b_icp, residual, t->icp_threads, t->icp_blocks);
                t->st.last_icp_error = sqrtf(residual[0]) / residual[1];
                t->st.last_icp_count = residual[1];
            }
            if(rgb)
                rgbStep(t->corresImg[i], sigmaVal, t->pointClouds[i], t->intr(i).fx, t->intr(i).fy, t->nextdIdx[i],
                        t->nextdIdy[i], t->sobel_scale, t->sumDataSE3, t->outDataSE3, A_rgb, b_rgb, t->rgb_threads,
                        t->rgb_blocks);

            double * lastA = t->st.last_A, * lastb = t->st.last_b, result[6];
            if(icp && rgb)
            {
                double w = icp_weight;
                for(int k = 0; k < 36; k++) lastA[k] = (double)A_rgb[k] + w * w * (double)A_icp[k];
                for(int k = 0; k < 6; k++) lastb[k] = (double)b_rgb[k] + w * (double)b_icp[k];
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
            efo_ldlt_solve_f64(lastA, lastb, 6, result);
            t->st.se3_iterations[i]++;

            double Rupd[9], Rtupd[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, newRt[16];
            efo_rodrigues(result + 3, Rupd);
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++) Rtupd[r * 4 + c] = Rupd[r * 3 + c];
                Rtupd[r * 4 + 3] = result[r];
            }
            for(int r = 0; r < 4; r++)
                for(int c = 0; c < 4; c++)
                {
                    double s = 0;
                    for(int k = 0; k < 4; k++) s += Rtupd[r * 4 + k] * resultRt[k * 4 + c];
                    newRt[r * 4 + c] = s;
                }
            memcpy(resultRt, newRt, sizeof(newRt));

            float oR[9], ot[3], iR[9], it[3];
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++) oR[r * 3 + c] = (float)resultRt[r * 4 + c];
                ot[r] = (float)resultRt[r * 4 + 3];
            }
            for(int r = 0; r < 3; r++)
                for(int c = 0; c < 3; c++) iR[r * 3 + c] = oR[c * 3 + r];
            for(int r = 0; r < 3; r++) it[r] = -(iR[r * 3] * ot[0] + iR[r * 3 + 1] * ot[1] + iR[r * 3 + 2] * ot[2]);
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++)
                    Rcurr[r * 3 + c] = Rprev[r * 3] * iR[c] + Rprev[r * 3 + 1] * iR[3 + c] + Rprev[r * 3 + 2] * iR[6 + c];
                tcurr[r] = Rprev[r * 3] * it[0] + Rprev[r * 3 + 1] * it[1] + Rprev[r * 3 + 2] * it[2] + tprev[r];
            }
        }
    }

    if(rgb)
    {
        float d[3] = {tcurr[0] - tprev[0], tcurr[1] - tprev[1], tcurr[2] - tprev[2]};
        if(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.3)
        {
            memcpy(Rcurr, Rprev, sizeof(Rcurr));
            memcpy(tcurr, tprev, sizeof(tcurr));
        }
    }

    if(so3)
        for(int i = 0; i < NUM_PYRS; i++) std::swap(t->lastNextImage[i], t->nextImage[i]);

    memcpy(trans, tcurr, sizeof(tcurr));
    memcpy(rot, Rcurr, sizeof(Rcurr));
    if(stats) *stats = t->st;
}

/* download an internal pyramid level into a dense host buffer; returns bytes written (0 = unknown name) */
EFR_API size_t efr_tracker_download(efr_tracker * t, const char * name, int level, void * host)
{
    if(level < 0 || level >= NUM_PYRS) return 0;
#define DL(nm, arr, T)                                                  \
    if(!strcmp(name, nm))                                               \
    {                                                                   \
        arr.download(host, arr.cols() * sizeof(T));                     \
        return (size_t)arr.rows() * arr.cols() * sizeof(T);             \
    }
    DL("vmap_curr", t->vmaps_curr[level], float)
    DL("nmap_curr", t->nmaps_curr[level], float)
    DL("vmap_g_prev", t->vmaps_g_prev[level], float)
    DL("nmap_g_prev", t->nmaps_g_prev[level], float)
    DL("last_depth", t->lastDepth[level], float)
    DL("next_depth", t->nextDepth[level], float)
    DL("last_image", t->lastImage[level], unsigned char)
    DL("next_image", t->nextImage[level], unsigned char)
    DL("last_next_image", t->lastNextImage[level], unsigned char)
    DL("dIdx", t->nextdIdx[level], short)
    DL("dIdy", t->nextdIdy[level], short)
    DL("depth_tmp", t->depth_tmp[level], unsigned short)
#undef DL
    return 0;
}
