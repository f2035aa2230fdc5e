// compat_harness.cu -- TEST INFRASTRUCTURE: compiles the two reference-side binding headers
//   include/compat/cudafuncs.cuh   (the 17 operator functions of Cuda/cudafuncs.cuh:64-177)
//   include/compat/RGBDOdometry.h  (class RGBDOdometry, Utils/RGBDOdometry.h:31-134)
// exactly as an ElasticFusion build would -- against the reference's OWN containers (Cuda/containers/*.hpp, device_memory.cpp)
// and types (Cuda/types.cuh), found on the include path under /root/reference, never copied -- and exposes a few C entry
// points so that tests/test_compat_gpu.py can run them on the GPU next to the reference kernels (oracle/_ref).
// GL and Eigen do not exist in this image: the textures are plain cudaArrays behind a mock GPUTexture (the header's
// documented customisation point) and <Eigen/Dense> is the stand-in under tests/compat/stub.
// Built by oracle/Makefile into oracle/_ref/libef_compat_test.so (git-ignored, travels to the GPU box).
#include <cuda_runtime.h>
#include <string.h>

#include <stdexcept>

// ---- the shim, fed from cudaArrays ----
struct GPUTexture
{
    cudaArray_t array;
};
#define EF_COMPAT_CUSTOM_TEXTURE_MAPPING
namespace ef_compat
{
struct MappedTexture
{
    cudaArray_t arr;
    explicit MappedTexture(GPUTexture * t) : arr(t->array) {}
};
} // namespace ef_compat
#include "compat/RGBDOdometry.h"

// ---- the operator functions over the reference's containers ----
#include "compat/cudafuncs.cuh"

#define EFC_API extern "C" __attribute__((visibility("default")))

static cudaArray_t make_array(const void * host, int w, int h, int bits, int channels, cudaChannelFormatKind kind)
{
    cudaChannelFormatDesc d = cudaCreateChannelDesc(bits, channels > 1 ? bits : 0, channels > 2 ? bits : 0, channels > 3 ? bits : 0, kind);
    cudaArray_t a = nullptr;
    if(cudaMallocArray(&a, &d, w, h) != cudaSuccess) return nullptr;
    const size_t row = (size_t)w * channels * bits / 8;
    cudaMemcpy2DToArray(a, 0, 0, host, row, row, h, cudaMemcpyHostToDevice);
    return a;
}

// One frameToModel step through class RGBDOdometry exactly as ElasticFusion::processFrame drives it (ElasticFusion.cpp:326,
// :343-368): initFirstRGB (when first), initICPModel, initRGBModel, initICP, initRGB, getIncrementalTransformation.
// `shim` persists across calls (state carried: the SO(3) image swap).  Returns 0, or -1 with the message in err.
struct ShimState
{
    RGBDOdometry * odom;
    int w, h;
};

EFC_API void * efc_shim_create(int w, int h, float cx, float cy, float fx, float fy)
{
    try
    {
        ShimState * s = new ShimState();
        s->odom = new RGBDOdometry(w, h, cx, cy, fx, fy); // the reference's default thresholds
        s->w = w;
        s->h = h;
        return s;
    }
    catch(const std::exception &)
    {
        return nullptr;
    }
}

EFC_API void efc_shim_destroy(void * p)
{
    ShimState * s = static_cast<ShimState *>(p);
    if(!s) return;
    delete s->odom;
    delete s;
}

EFC_API int efc_shim_frame(void * p, int first, const float * v4, const float * n4, const unsigned char * model_rgba, const unsigned short * depth,
                           const unsigned char * rgba, const float * pose16_rowmajor, int rgb_only, float icp_weight, int pyramid, int fast_odom,
                           int so3, float * trans3, float * rot9, float * stats6, double * lastA36, double * lastb6, double * cov36, char * err,
                           int err_len)
{
    ShimState * s = static_cast<ShimState *>(p);
    GPUTexture tv{make_array(v4, s->w, s->h, 32, 4, cudaChannelFormatKindFloat)}, tn{make_array(n4, s->w, s->h, 32, 4, cudaChannelFormatKindFloat)};
    GPUTexture tm{make_array(model_rgba, s->w, s->h, 8, 4, cudaChannelFormatKindUnsigned)}, tc{make_array(rgba, s->w, s->h, 8, 4, cudaChannelFormatKindUnsigned)};
    GPUTexture td{make_array(depth, s->w, s->h, 16, 1, cudaChannelFormatKindUnsigned)};
    int rc = 0;
    try
    {
        Eigen::Matrix4f pose;
        for(int r = 0; r < 4; r++)
            for(int c = 0; c < 4; c++) pose(r, c) = pose16_rowmajor[r * 4 + c];
        RGBDOdometry & o = *s->odom;
        if(first) o.initFirstRGB(&tm);
        o.initICPModel(&tv, &tn, 20.0f, pose);
        o.initRGBModel(&tm);
        o.initICP(&td, 20.0f);
        o.initRGB(&tc);
        Eigen::Vector3f trans;
        Eigen::Matrix<float, 3, 3, Eigen::RowMajor> rot;
        for(int r = 0; r < 3; r++)
        {
            trans(r, 0) = pose(r, 3);
            for(int c = 0; c < 3; c++) rot(r, c) = pose(r, c);
        }
        o.getIncrementalTransformation(trans, rot, rgb_only != 0, icp_weight, pyramid != 0, fast_odom != 0, so3 != 0);
        memcpy(trans3, trans.data(), 12);
        memcpy(rot9, rot.data(), 36);
        const float st[6] = {o.lastICPError, o.lastICPCount, o.lastRGBError, o.lastRGBCount, o.lastSO3Error, o.lastSO3Count};
        memcpy(stats6, st, sizeof(st));
        memcpy(lastA36, o.lastA.data(), 36 * sizeof(double));
        memcpy(lastb6, o.lastb.data(), 6 * sizeof(double));
        Eigen::MatrixXd cov = o.getCovariance();
        for(int i = 0; i < 6; i++)
            for(int j = 0; j < 6; j++) cov36[i * 6 + j] = cov(i, j);
    }
    catch(const std::exception & e)
    {
        strncpy(err, e.what(), err_len - 1);
        err[err_len - 1] = 0;
        rc = -1;
    }
    cudaDeviceSynchronize();
    for(GPUTexture * t : {&tv, &tn, &tm, &tc, &td}) cudaFreeArray(t->array);
    return rc;
}

// The drop-in rate in C++: `iters` frameToModel steps of ONE frame pair through class RGBDOdometry exactly as efc_shim_frame drives
// it -- textures resident (cudaArrays, as the reference's GL textures are), five init calls + getIncrementalTransformation, the pose
// read back every frame -- timed with the host clock around the loop.  defer != 0: EF_OPT_DEFER_BUILD, the setting INTEGRATION.md
// recommends for the shim (the init calls copy the texels and record; one builder launch at the solve).  Returns frames/s, < 0 on error.
#include <chrono>
EFC_API double efc_shim_bench(void * p, const float * v4, const float * n4, const unsigned char * model_rgba, const unsigned short * depth,
                              const unsigned char * rgba, const float * pose16_rowmajor, int iters, int defer)
{
    ShimState * s = static_cast<ShimState *>(p);
    GPUTexture tv{make_array(v4, s->w, s->h, 32, 4, cudaChannelFormatKindFloat)}, tn{make_array(n4, s->w, s->h, 32, 4, cudaChannelFormatKindFloat)};
    GPUTexture tm{make_array(model_rgba, s->w, s->h, 8, 4, cudaChannelFormatKindUnsigned)}, tc{make_array(rgba, s->w, s->h, 8, 4, cudaChannelFormatKindUnsigned)};
    GPUTexture td{make_array(depth, s->w, s->h, 16, 1, cudaChannelFormatKindUnsigned)};
    double fps = -1.0;
    try
    {
        Eigen::Matrix4f pose;
        for(int r = 0; r < 4; r++)
            for(int c = 0; c < 4; c++) pose(r, c) = pose16_rowmajor[r * 4 + c];
        RGBDOdometry & o = *s->odom;
        ef_tracker_set_option(o.tracker(), EF_OPT_DEFER_BUILD, defer ? 1 : 0);
        auto frame = [&]() {
            o.initICPModel(&tv, &tn, 20.0f, pose);
            o.initRGBModel(&tm);
            o.initICP(&td, 20.0f);
            o.initRGB(&tc);
            Eigen::Vector3f trans;
            Eigen::Matrix<float, 3, 3, Eigen::RowMajor> rot;
            for(int r = 0; r < 3; r++)
            {
                trans(r, 0) = pose(r, 3);
                for(int c = 0; c < 3; c++) rot(r, c) = pose(r, c);
            }
            o.getIncrementalTransformation(trans, rot, false, 10.0f, true, false, false);
            return trans(0, 0);
        };
        float sink = 0.f;
        for(int i = 0; i < 10; i++) sink += frame();
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        for(int i = 0; i < iters; i++) sink += frame();
        cudaDeviceSynchronize();
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fps = (sink == 12345.678f) ? 0.0 : iters / sec;
        ef_tracker_set_option(o.tracker(), EF_OPT_DEFER_BUILD, 0);
    }
    catch(const std::exception &)
    {
        fps = -1.0;
    }
    for(GPUTexture * t : {&tv, &tn, &tm, &tc, &td}) cudaFreeArray(t->array);
    return fps;
}

// ---- operator functions: the depth pyramid + vertex / normal maps, the intensity pyramid + derivatives, one icpStep ----
EFC_API int efc_ops_depth_chain(const unsigned short * depth, int rows, int cols, float fx, float fy, float cx, float cy, float cutoff,
                                unsigned short * depth1 /*rows/2 x cols/2*/, float * vmap1 /*3*rows/2 x cols/2*/, float * nmap1, char * err, int err_len)
{
    try
    {
        DeviceArray2D<unsigned short> d0(rows, cols), d1;
        d0.upload(depth, (size_t)cols * 2, rows, cols);
        pyrDown(d0, d1);
        DeviceArray2D<float> v1, n1;
        CameraModel intr(fx, fy, cx, cy);
        createVMap(intr(1), d1, v1, cutoff);
        createNMap(v1, n1);
        d1.download(depth1, (size_t)(cols / 2) * 2);
        v1.download(vmap1, (size_t)(cols / 2) * 4);
        n1.download(nmap1, (size_t)(cols / 2) * 4);
        return 0;
    }
    catch(const std::exception & e)
    {
        strncpy(err, e.what(), err_len - 1);
        err[err_len - 1] = 0;
        return -1;
    }
}

EFC_API int efc_ops_image_chain(const unsigned char * rgba, int rows, int cols, unsigned char * img0, unsigned char * img1, short * dx1, short * dy1,
                                char * err, int err_len)
{
    cudaArray_t arr = make_array(rgba, cols, rows, 8, 4, cudaChannelFormatKindUnsigned);
    int rc = 0;
    try
    {
        DeviceArray2D<unsigned char> i0(rows, cols), i1;
        imageBGRToIntensity(arr, i0);
        pyrDownUcharGauss(i0, i1);
        DeviceArray2D<short> gx(rows / 2, cols / 2), gy(rows / 2, cols / 2);
        computeDerivativeImages(i1, gx, gy);
        i0.download(img0, cols);
        i1.download(img1, cols / 2);
        gx.download(dx1, (size_t)(cols / 2) * 2);
        gy.download(dy1, (size_t)(cols / 2) * 2);
    }
    catch(const std::exception & e)
    {
        strncpy(err, e.what(), err_len - 1);
        err[err_len - 1] = 0;
        rc = -1;
    }
    cudaFreeArray(arr);
    return rc;
}

// copyMaps + resize + tranformMaps (initICPModel's chain) then icpStep of level 1 against the current maps of a depth image
EFC_API int efc_ops_icp(const float * v4, const float * n4, const unsigned short * depth, int rows, int cols, float fx, float fy, float cx, float cy,
                        const float * pose16_rowmajor, float * A36, float * b6, float * residual2, char * err, int err_len)
{
    try
    {
        CameraModel intr(fx, fy, cx, cy);
        DeviceArray<float> vt((size_t)rows * cols * 4), nt((size_t)rows * cols * 4);
        vt.upload(v4, (size_t)rows * cols * 4);
        nt.upload(n4, (size_t)rows * cols * 4);
        DeviceArray2D<float> vp0(rows * 3, cols), np0(rows * 3, cols), vp1, np1;
        copyMaps(vt, nt, vp0, np0);
        resizeVMap(vp0, vp1);
        resizeNMap(np0, np1);
        mat33 R;
        float3 t;
        for(int r = 0; r < 3; r++) R.data[r] = make_float3(pose16_rowmajor[r * 4], pose16_rowmajor[r * 4 + 1], pose16_rowmajor[r * 4 + 2]);
        t = make_float3(pose16_rowmajor[3], pose16_rowmajor[7], pose16_rowmajor[11]);
        tranformMaps(vp1, np1, R, t, vp1, np1);
        DeviceArray2D<unsigned short> d0(rows, cols), d1;
        d0.upload(depth, (size_t)cols * 2, rows, cols);
        pyrDown(d0, d1);
        DeviceArray2D<float> vc1, nc1;
        createVMap(intr(1), d1, vc1, 20.0f);
        createNMap(vc1, nc1);
        // Rprev^-1 = R^T for a rotation
        mat33 Rinv;
        Rinv.data[0] = make_float3(R.data[0].x, R.data[1].x, R.data[2].x);
        Rinv.data[1] = make_float3(R.data[0].y, R.data[1].y, R.data[2].y);
        Rinv.data[2] = make_float3(R.data[0].z, R.data[1].z, R.data[2].z);
        DeviceArray<JtJJtrSE3> sum(1024), out(1);
        icpStep(R, t, vc1, nc1, Rinv, t, intr(1), vp1, np1, 0.10f, sinf(20.f * 3.14159254f / 180.f), sum, out, A36, b6, residual2, 128, 112);
        return 0;
    }
    catch(const std::exception & e)
    {
        strncpy(err, e.what(), err_len - 1);
        err[err_len - 1] = 0;
        return -1;
    }
}
