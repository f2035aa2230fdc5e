/*
 * compat/RGBDOdometry.h -- drop-in replacement of elasticfusionpublic/Core/src/Utils/RGBDOdometry.{h,cpp}.
 *
 * class RGBDOdometry with the reference's constructor, methods, default arguments and public fields
 * (Utils/RGBDOdometry.h:31-73), implemented header-only over the C ABI of libef_track.so (include/ef_track.h).
 * ElasticFusion::processFrame (ElasticFusion.cpp:326, :343-368, :531-548), Ferns::findFrame (Ferns.cpp:576-592) and
 * GPUTest.cpp:215-278 compile against it unchanged; Utils/RGBDOdometry.cpp, Cuda/cudafuncs.cu and Cuda/reduce.cu drop out
 * of the build.  Eigen and GPUTexture appear only here, never in the library.
 *
 * The GL textures are mapped exactly where the reference maps them -- for the duration of an init* call
 * (RGBDOdometry.cpp:120-128) -- and handed over as cudaArray_t (ef_init_*_array); the handle copies the texels into its
 * own staging buffer on its stream, so the unmap waits for that copy only (one event, not the reference's
 * cudaDeviceSynchronize).  With EF_OPT_DEFER_BUILD (set below) the four init* calls of a frame only stage their inputs
 * and getIncrementalTransformation builds every pyramid with one launch before the persistent tracker kernel:
 * 2 kernel launches per tracked frame.
 *
 * Texture access is the one customisation point: define EF_COMPAT_CUSTOM_TEXTURE_MAPPING and provide
 * ef_compat::MappedTexture (constructor from GPUTexture*, member `cudaArray_t arr`) before including this header to
 * feed the shim from something other than a registered GL texture (tests/compat/compat_harness.cu does, from plain
 * cudaArrays, because this image has no GL).
 */
#ifndef EF_COMPAT_RGBDODOMETRY_H_
#define EF_COMPAT_RGBDODOMETRY_H_

#include <Eigen/Dense>
#include <cuda_runtime_api.h>

#include <cmath>
#include <stdexcept>
#include <string>

#include "../ef_track.h"
#ifdef EF_COMPAT_WITH_STOPWATCH
#include "Stopwatch.h" /* Core/src/Utils/Stopwatch.h of the including project */
#endif

#ifndef EF_COMPAT_CUSTOM_TEXTURE_MAPPING
#include "../GPUTexture.h" /* Core/src/GPUTexture.h: `cudaGraphicsResource * cudaRes`, registered at construction (GPUTexture.cpp:40-47) */

namespace ef_compat
{
/* cudaGraphicsMapResources ... cudaGraphicsUnmapResources around one init* call (RGBDOdometry.cpp:120-128) */
struct MappedTexture
{
    cudaGraphicsResource * res;
    cudaArray_t arr;
    explicit MappedTexture(GPUTexture * t) : res(t->cudaRes), arr(nullptr)
    {
        if(cudaGraphicsMapResources(1, &res) != cudaSuccess || cudaGraphicsSubResourceGetMappedArray(&arr, res, 0, 0) != cudaSuccess)
            throw std::runtime_error("RGBDOdometry: mapping the GL texture failed");
    }
    ~MappedTexture() { cudaGraphicsUnmapResources(1, &res); }
    MappedTexture(const MappedTexture &) = delete;
    MappedTexture & operator=(const MappedTexture &) = delete;
};
} // namespace ef_compat
#endif

class RGBDOdometry
{
  public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    RGBDOdometry(int width, int height, float cx, float cy, float fx, float fy, float distThresh = 0.10f,
                 float angleThresh = sin(20.f * 3.14159254f / 180.f))
     : lastICPError(0), lastICPCount(width * height), lastRGBError(0), lastRGBCount(width * height), lastSO3Error(0), lastSO3Count(width * height),
       lastA(Eigen::Matrix<double, 6, 6, Eigen::RowMajor>::Zero()), lastb(Eigen::Matrix<double, 6, 1>::Zero()), handle(nullptr), staged(nullptr)
    {
        const int rc = ef_tracker_create(width, height, cx, cy, fx, fy, distThresh, angleThresh, /*stream: the handle's own*/ nullptr, &handle);
        if(rc) throw std::runtime_error("RGBDOdometry: ef_tracker_create failed with code " + std::to_string(rc));
        /* the staged inputs live in handle-owned buffers until the solve, so deferring the pyramid build is safe here */
        ef_tracker_set_option(handle, EF_OPT_DEFER_BUILD, 1);
        if(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming) != cudaSuccess) staged = nullptr;
    }

    virtual ~RGBDOdometry()
    {
        if(staged) cudaEventDestroy(staged);
        ef_tracker_destroy(handle);
    }

    RGBDOdometry(const RGBDOdometry &) = delete;
    RGBDOdometry & operator=(const RGBDOdometry &) = delete;

    /* RGBDOdometry.cpp:118-142 */
    void initICP(GPUTexture * filteredDepth, const float depthCutoff)
    {
        ef_compat::MappedTexture d(filteredDepth);
        check(ef_init_icp_depth_array(handle, d.arr, depthCutoff));
        texels_copied();
    }

    /* RGBDOdometry.cpp:144-167 */
    void initICP(GPUTexture * predictedVertices, GPUTexture * predictedNormals, const float depthCutoff)
    {
        ef_compat::MappedTexture v(predictedVertices), n(predictedNormals);
        check(ef_init_icp_maps_array(handle, v.arr, n.arr, depthCutoff));
        texels_copied();
    }

    /* RGBDOdometry.cpp:169-206 */
    void initICPModel(GPUTexture * predictedVertices, GPUTexture * predictedNormals, const float depthCutoff, const Eigen::Matrix4f & modelPose)
    {
        float pose[16]; /* row-major, whatever the storage order of the caller's matrix */
        for(int r = 0; r < 4; r++)
            for(int c = 0; c < 4; c++) pose[r * 4 + c] = modelPose(r, c);
        ef_compat::MappedTexture v(predictedVertices), n(predictedNormals);
        check(ef_init_icp_model_array(handle, v.arr, n.arr, depthCutoff, pose));
        texels_copied();
    }

    /* RGBDOdometry.cpp:243-247, :237-241, :249-265 */
    void initRGB(GPUTexture * rgb)
    {
        ef_compat::MappedTexture r(rgb);
        check(ef_init_rgb_array(handle, r.arr));
        texels_copied();
    }
    void initRGBModel(GPUTexture * rgb)
    {
        ef_compat::MappedTexture r(rgb);
        check(ef_init_rgb_model_array(handle, r.arr));
        texels_copied();
    }
    void initFirstRGB(GPUTexture * rgb)
    {
        ef_compat::MappedTexture r(rgb);
        check(ef_init_first_rgb_array(handle, r.arr));
        texels_copied();
    }

    /* RGBDOdometry.cpp:267-603 */
    void getIncrementalTransformation(Eigen::Vector3f & trans, Eigen::Matrix<float, 3, 3, Eigen::RowMajor> & rot, const bool & rgbOnly,
                                      const float & icpWeight, const bool & pyramid, const bool & fastOdom, const bool & so3)
    {
        ef_track_stats st;
        check(ef_get_incremental_transformation(handle, trans.data(), rot.data(), rgbOnly, icpWeight, pyramid, fastOdom, so3, &st));
        lastICPError = st.last_icp_error;
        lastICPCount = st.last_icp_count;
        lastRGBError = st.last_rgb_error;
        lastRGBCount = st.last_rgb_count;
        lastSO3Error = st.last_so3_error;
        lastSO3Count = st.last_so3_count;
        for(int i = 0; i < 6; i++)
        {
            for(int j = 0; j < 6; j++) lastA(i, j) = st.last_A[i * 6 + j];
            lastb(i, 0) = st.last_b[i];
        }
#ifdef EF_COMPAT_WITH_STOPWATCH
        /* the reference's TICK / TOCK keys of this function (RGBDOdometry.cpp:333-538), for GPUTest.cpp:283-286 and the GUI's plots;
         * Stopwatch counts microseconds (Utils/Stopwatch.h:100-106) */
        ef_stage_times tm;
        if(ef_tracker_stage_times(handle, &tm) == 0 && tm.solve_mode == EF_SOLVE_HOST)
        {
            Stopwatch::getInstance().addStopwatchTiming("so3Step", (unsigned long long)(tm.so3_step_ms * 1000.f));
            Stopwatch::getInstance().addStopwatchTiming("computeRgbResidual", (unsigned long long)(tm.rgb_residual_ms * 1000.f));
            Stopwatch::getInstance().addStopwatchTiming("icpStep", (unsigned long long)(tm.icp_step_ms * 1000.f));
            Stopwatch::getInstance().addStopwatchTiming("rgbStep", (unsigned long long)(tm.rgb_step_ms * 1000.f));
        }
#endif
    }

    /* RGBDOdometry.cpp:605-608 */
    Eigen::MatrixXd getCovariance()
    {
        double c[36];
        check(ef_get_covariance(handle, c));
        Eigen::MatrixXd cov(6, 6);
        for(int i = 0; i < 6; i++)
            for(int j = 0; j < 6; j++) cov(i, j) = c[i * 6 + j];
        return cov;
    }

    float lastICPError;
    float lastICPCount;
    float lastRGBError;
    float lastRGBCount;
    float lastSO3Error;
    float lastSO3Count;

    Eigen::Matrix<double, 6, 6, Eigen::RowMajor> lastA;
    Eigen::Matrix<double, 6, 1> lastb;

    /* not in the reference: the handle, for callers that want the library's options (EF_OPT_*) or its stream */
    ef_tracker * tracker() { return handle; }

  private:
    /* the texture may be unmapped (and rendered to again by GL) once the staging copy out of its array has run: wait for
     * exactly that copy; the pyramid builders and the solve stay asynchronous behind it */
    void texels_copied()
    {
        cudaStream_t s = static_cast<cudaStream_t>(ef_tracker_stream(handle));
        if(staged && cudaEventRecord(staged, s) == cudaSuccess && cudaEventSynchronize(staged) == cudaSuccess) return;
        check(ef_tracker_synchronize(handle));
    }
    void check(int rc)
    {
        if(rc) throw std::runtime_error(std::string("RGBDOdometry: ") + ef_last_error(handle) + " (code " + std::to_string(rc) + ")");
    }

    ef_tracker * handle;
    cudaEvent_t staged;
};

#endif /* EF_COMPAT_RGBDODOMETRY_H_ */
