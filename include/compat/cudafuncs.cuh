/*
 * compat/cudafuncs.cuh -- source-compatible replacement of elasticfusionpublic/Core/src/Cuda/cudafuncs.cuh:64-177.
 *
 * The 17 free operator functions ElasticFusion's host code calls (RGBDOdometry.cpp, Ferns.cpp, GPUTest.cpp), with the
 * reference's names, argument lists and argument meaning, each forwarding to its C-ABI twin in libef_track.so
 * (include/ef_track.h, Tier 2).  Nothing of Cuda/cudafuncs.cu or Cuda/reduce.cu is compiled any more.
 *
 * The CONTAINERS stay the caller's: like the reference header (cudafuncs.cuh:61-62) this one includes
 * "containers/device_array.hpp" and "types.cuh" from the including project's Cuda directory -- DeviceArray / DeviceArray2D
 * (ptr(), step(), rows(), cols()), mat33, CameraModel, DataTerm, JtJJtrSE3 / JtJJtrSO3 are used only through their public
 * members.  Put this file's directory before Core/src/Cuda on the include path, or replace Cuda/cudafuncs.cuh by it.
 *
 * Differences from the reference, all on purpose:
 *   - `threads` / `blocks` (GPUConfig) are accepted and ignored: launch shapes follow the image size and the SM count;
 *   - `sum` / `out` (the 1024-row reduction buffers) are accepted and ignored: the reductions use a scratch block that
 *     this header allocates once per host thread (ef_op_scratch_bytes());
 *   - a CUDA failure throws std::runtime_error instead of printing and calling exit(0) (Cuda/convenience.cuh:64-71);
 *   - everything runs on the legacy default stream like the reference; functions that return host results
 *     synchronise it, the others are asynchronous (the reference adds a cudaDeviceSynchronize to most of them).
 */
#ifndef EF_COMPAT_CUDAFUNCS_CUH_
#define EF_COMPAT_CUDAFUNCS_CUH_

#include <cuda_runtime_api.h>
#include <vector_types.h>

#include <stdexcept>
#include <string>

#include "containers/device_array.hpp" /* the caller's containers (Cuda/containers/) */
#include "types.cuh"                   /* the caller's mat33, CameraModel, DataTerm, JtJJtrSE3, JtJJtrSO3 */

#include "../ef_track.h"

namespace ef_compat
{

inline void check(int rc, const char * what)
{
    if(rc == 0) return;
    std::string msg = std::string(what) + " failed: ";
    msg += rc > 0 ? cudaGetErrorString((cudaError_t)rc) : ("EF error " + std::to_string(rc)).c_str();
    throw std::runtime_error(msg);
}

/* one device block per host thread: reduction scratch, then a linear staging image for imageBGRToIntensity */
struct Workspace
{
    void * scratch = nullptr;
    void * stage = nullptr;
    size_t stage_bytes = 0;
    ~Workspace()
    {
        if(scratch) cudaFree(scratch);
        if(stage) cudaFree(stage);
    }
    void * reduction()
    {
        if(!scratch)
        {
            check((int)cudaMalloc(&scratch, ef_op_scratch_bytes()), "cudaMalloc(reduction scratch)");
            check((int)cudaMemset(scratch, 0, ef_op_scratch_bytes()), "cudaMemset(reduction scratch)");
        }
        return scratch;
    }
    void * staging(size_t bytes)
    {
        if(bytes > stage_bytes)
        {
            if(stage) cudaFree(stage);
            stage = nullptr;
            check((int)cudaMalloc(&stage, bytes), "cudaMalloc(staging image)");
            stage_bytes = bytes;
        }
        return stage;
    }
};

inline Workspace & workspace()
{
    static thread_local Workspace w;
    return w;
}

inline const float * m33(const mat33 & m) { return &m.data[0].x; } /* 3 x float3 rows = 9 floats, row-major (types.cuh:61-73) */
inline const float * f3(const float3 & v) { return &v.x; }

} // namespace ef_compat

/* ---- reduce.cu ---- */

/* cudafuncs.cuh:64-81 / reduce.cu:390-490 */
inline void icpStep(const mat33 & Rcurr, const float3 & tcurr, const DeviceArray2D<float> & vmap_curr, const DeviceArray2D<float> & nmap_curr,
                    const mat33 & Rprev_inv, const float3 & tprev, const CameraModel & intr, const DeviceArray2D<float> & vmap_g_prev,
                    const DeviceArray2D<float> & nmap_g_prev, float distThres, float angleThres, DeviceArray<JtJJtrSE3> & /*sum*/,
                    DeviceArray<JtJJtrSE3> & /*out*/, float * matrixA_host, float * vectorB_host, float * residual_host, int /*threads*/,
                    int /*blocks*/)
{
    const int cols = vmap_curr.cols(), rows = vmap_curr.rows() / 3; /* reduce.cu:408-409 */
    if(nmap_curr.step() != vmap_curr.step() || vmap_g_prev.step() != vmap_curr.step() || nmap_g_prev.step() != vmap_curr.step())
        throw std::runtime_error("icpStep: the four maps must share one pitch");
    ef_compat::check(ef_op_icp_step(ef_compat::m33(Rcurr), ef_compat::f3(tcurr), vmap_curr.ptr(), nmap_curr.ptr(), ef_compat::m33(Rprev_inv),
                                    ef_compat::f3(tprev), intr.fx, intr.fy, intr.cx, intr.cy, vmap_g_prev.ptr(), nmap_g_prev.ptr(), vmap_curr.step(),
                                    distThres, angleThres, rows, cols, ef_compat::workspace().reduction(), matrixA_host, vectorB_host, residual_host,
                                    nullptr),
                     "icpStep");
}

/* cudafuncs.cuh:83-96 / reduce.cu:597-678 */
inline void rgbStep(const DeviceArray2D<DataTerm> & corresImg, const float & sigma, const DeviceArray2D<float3> & cloud, const float & fx,
                    const float & fy, const DeviceArray2D<short> & dIdx, const DeviceArray2D<short> & dIdy, const float & sobelScale,
                    DeviceArray<JtJJtrSE3> & /*sum*/, DeviceArray<JtJJtrSE3> & /*out*/, float * matrixA_host, float * vectorB_host, int /*threads*/,
                    int /*blocks*/)
{
    ef_compat::check(ef_op_rgb_step(corresImg.ptr(), sigma, reinterpret_cast<const float *>(cloud.ptr()), cloud.step(), fx, fy, dIdx.ptr(), dIdy.ptr(),
                                    dIdx.step(), sobelScale, dIdx.rows(), dIdx.cols(), ef_compat::workspace().reduction(), matrixA_host, vectorB_host,
                                    nullptr),
                     "rgbStep");
}

/* cudafuncs.cuh:98-109 / reduce.cu:1058-1141 */
inline void so3Step(const DeviceArray2D<unsigned char> & lastImage, const DeviceArray2D<unsigned char> & nextImage, const mat33 & imageBasis,
                    const mat33 & kinv, const mat33 & krlr, DeviceArray<JtJJtrSO3> & /*sum*/, DeviceArray<JtJJtrSO3> & /*out*/, float * matrixA_host,
                    float * vectorB_host, float * residual_host, int /*threads*/, int /*blocks*/)
{
    ef_compat::check(ef_op_so3_step(lastImage.ptr(), nextImage.ptr(), lastImage.step(), ef_compat::m33(imageBasis), ef_compat::m33(kinv),
                                    ef_compat::m33(krlr), lastImage.rows(), lastImage.cols(), ef_compat::workspace().reduction(), matrixA_host,
                                    vectorB_host, residual_host, nullptr),
                     "so3Step");
}

/* cudafuncs.cuh:111-126 / reduce.cu:878-936 */
inline void computeRgbResidual(const float & minScale, const DeviceArray2D<short> & dIdx, const DeviceArray2D<short> & dIdy,
                               const DeviceArray2D<float> & lastDepth, const DeviceArray2D<float> & nextDepth,
                               const DeviceArray2D<unsigned char> & lastImage, const DeviceArray2D<unsigned char> & nextImage,
                               DeviceArray2D<DataTerm> & corresImg, DeviceArray<int2> & /*sumResidual*/, const float maxDepthDelta, const float3 & kt,
                               const mat33 & krkinv, int & sigmaSum, int & count, int /*threads*/, int /*blocks*/)
{
    if(dIdy.step() != dIdx.step() || nextDepth.step() != lastDepth.step() || nextImage.step() != lastImage.step())
        throw std::runtime_error("computeRgbResidual: image pairs must share their pitch");
    ef_compat::check(ef_op_rgb_residual(minScale, dIdx.ptr(), dIdy.ptr(), dIdx.step(), lastDepth.ptr(), nextDepth.ptr(), lastDepth.step(), lastImage.ptr(),
                                        nextImage.ptr(), lastImage.step(), corresImg.ptr(), maxDepthDelta, ef_compat::f3(kt), ef_compat::m33(krkinv),
                                        nextImage.rows(), nextImage.cols(), ef_compat::workspace().reduction(), &sigmaSum, &count, nullptr),
                     "computeRgbResidual");
}

/* ---- cudafuncs.cu ---- */

/* cudafuncs.cuh:128-131 / cudafuncs.cu:134-149 */
inline void createVMap(const CameraModel & intr, const DeviceArray2D<unsigned short> & depth, DeviceArray2D<float> & vmap, const float depthCutoff)
{
    vmap.create(depth.rows() * 3, depth.cols());
    ef_compat::check(ef_op_create_vmap(depth.ptr(), depth.step(), depth.rows(), depth.cols(), intr.fx, intr.fy, intr.cx, intr.cy, depthCutoff, vmap.ptr(),
                                       vmap.step(), nullptr),
                     "createVMap");
}

/* cudafuncs.cuh:133-134 / cudafuncs.cu:190-204 */
inline void createNMap(const DeviceArray2D<float> & vmap, DeviceArray2D<float> & nmap)
{
    nmap.create(vmap.rows(), vmap.cols());
    if(nmap.step() != vmap.step()) throw std::runtime_error("createNMap: pitch mismatch");
    ef_compat::check(ef_op_create_nmap(vmap.ptr(), vmap.step(), vmap.rows() / 3, vmap.cols(), nmap.ptr(), nmap.step(), nullptr), "createNMap");
}

/* cudafuncs.cuh:136-141 / cudafuncs.cu:250-268 (sic: "tranformMaps") */
inline void tranformMaps(const DeviceArray2D<float> & vmap_src, const DeviceArray2D<float> & nmap_src, const mat33 & Rmat, const float3 & tvec,
                         DeviceArray2D<float> & vmap_dst, DeviceArray2D<float> & nmap_dst)
{
    const int cols = vmap_src.cols(), rows = vmap_src.rows() / 3;
    vmap_dst.create(rows * 3, cols);
    nmap_dst.create(rows * 3, cols);
    ef_compat::check(ef_op_transform_maps(vmap_src.ptr(), nmap_src.ptr(), vmap_src.step(), rows, cols, ef_compat::m33(Rmat), ef_compat::f3(tvec),
                                          vmap_dst.ptr(), nmap_dst.ptr(), vmap_dst.step(), nullptr),
                     "tranformMaps");
}

/* cudafuncs.cuh:143-146 / cudafuncs.cu:312-330: RGBA32F textures (linear, 4 floats per pixel) -> 3-plane maps */
inline void copyMaps(const DeviceArray<float> & vmap_src, const DeviceArray<float> & nmap_src, DeviceArray2D<float> & vmap_dst,
                     DeviceArray2D<float> & nmap_dst)
{
    const int cols = vmap_dst.cols(), rows = vmap_dst.rows() / 3;
    ef_compat::check(ef_op_copy_maps(vmap_src.ptr(), nmap_src.ptr(), rows, cols, vmap_dst.ptr(), nmap_dst.ptr(), vmap_dst.step(), nullptr), "copyMaps");
}

/* cudafuncs.cuh:148-152 / cudafuncs.cu:420-444 */
inline void resizeVMap(const DeviceArray2D<float> & input, DeviceArray2D<float> & output)
{
    const int in_cols = input.cols(), in_rows = input.rows() / 3;
    output.create((in_rows / 2) * 3, in_cols / 2);
    ef_compat::check(ef_op_resize_map(input.ptr(), input.step(), in_rows, in_cols, output.ptr(), output.step(), 0, nullptr), "resizeVMap");
}

inline void resizeNMap(const DeviceArray2D<float> & input, DeviceArray2D<float> & output)
{
    const int in_cols = input.cols(), in_rows = input.rows() / 3;
    output.create((in_rows / 2) * 3, in_cols / 2);
    ef_compat::check(ef_op_resize_map(input.ptr(), input.step(), in_rows, in_cols, output.ptr(), output.step(), 1, nullptr), "resizeNMap");
}

/* cudafuncs.cuh:154-155 / cudafuncs.cu:565-577: the reference samples the mapped GL texture through a texture reference;
 * here the array is copied into a linear image first (same texels, same arithmetic) */
inline void imageBGRToIntensity(cudaArray * cuArr, DeviceArray2D<unsigned char> & dst)
{
    const size_t row = (size_t)dst.cols() * 4;
    void * lin = ef_compat::workspace().staging(row * dst.rows());
    ef_compat::check((int)cudaMemcpy2DFromArrayAsync(lin, row, cuArr, 0, 0, row, dst.rows(), cudaMemcpyDeviceToDevice, nullptr), "imageBGRToIntensity(copy)");
    ef_compat::check(ef_op_bgr_to_intensity(static_cast<const uint8_t *>(lin), row, dst.rows(), dst.cols(), dst.ptr(), dst.step(), nullptr),
                     "imageBGRToIntensity");
}

/* cudafuncs.cuh:157-159 / cudafuncs.cu:540-546 */
inline void verticesToDepth(DeviceArray<float> & vmap_src, DeviceArray2D<float> & dst, float cutOff)
{
    ef_compat::check(ef_op_vertices_to_depth(vmap_src.ptr(), dst.rows(), dst.cols(), cutOff, dst.ptr(), dst.step(), nullptr), "verticesToDepth");
}

/* cudafuncs.cuh:161-164 / cudafuncs.cu:661-674 */
inline void projectToPointCloud(const DeviceArray2D<float> & depth, const DeviceArray2D<float3> & cloud, CameraModel & intrinsics, const int & level)
{
    ef_compat::check(ef_op_project_point_cloud(depth.ptr(), depth.step(), depth.rows(), depth.cols(), intrinsics.fx, intrinsics.fy, intrinsics.cx,
                                               intrinsics.cy, level, reinterpret_cast<float *>(const_cast<float3 *>(cloud.ptr())), cloud.step(), nullptr),
                     "projectToPointCloud");
}

/* cudafuncs.cuh:166-167 / cudafuncs.cu:97-107 */
inline void pyrDown(const DeviceArray2D<unsigned short> & src, DeviceArray2D<unsigned short> & dst)
{
    dst.create(src.rows() / 2, src.cols() / 2);
    ef_compat::check(ef_op_pyr_down_u16(src.ptr(), src.step(), src.rows(), src.cols(), dst.ptr(), dst.step(), nullptr), "pyrDown");
}

/* cudafuncs.cuh:169-170 / cudafuncs.cu:446-468 */
inline void pyrDownGaussF(const DeviceArray2D<float> & src, DeviceArray2D<float> & dst)
{
    dst.create(src.rows() / 2, src.cols() / 2);
    ef_compat::check(ef_op_pyr_down_gauss_f32(src.ptr(), src.step(), src.rows(), src.cols(), dst.ptr(), dst.step(), nullptr), "pyrDownGaussF");
}

/* cudafuncs.cuh:172-173 / cudafuncs.cu:502-524 */
inline void pyrDownUcharGauss(const DeviceArray2D<unsigned char> & src, DeviceArray2D<unsigned char> & dst)
{
    dst.create(src.rows() / 2, src.cols() / 2);
    ef_compat::check(ef_op_pyr_down_gauss_u8(src.ptr(), src.step(), src.rows(), src.cols(), dst.ptr(), dst.step(), nullptr), "pyrDownUcharGauss");
}

/* cudafuncs.cuh:175-177 / cudafuncs.cu:608-639 */
inline void computeDerivativeImages(DeviceArray2D<unsigned char> & src, DeviceArray2D<short> & dx, DeviceArray2D<short> & dy)
{
    if(dy.step() != dx.step()) throw std::runtime_error("computeDerivativeImages: dx and dy must share their pitch");
    ef_compat::check(ef_op_derivative_images(src.ptr(), src.step(), src.rows(), src.cols(), dx.ptr(), dy.ptr(), dx.step(), nullptr),
                     "computeDerivativeImages");
}

#endif /* EF_COMPAT_CUDAFUNCS_CUH_ */
