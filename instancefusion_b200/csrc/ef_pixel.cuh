// ef_pixel.cuh -- per-pixel math of the four association steps, shared by the stand-alone
// operator kernels (ef_ops_reduce.cu) and the persistent tracker kernel (ef_track_kernel.cu).
//
// Citations are to elasticfusionpublic/Core/src/Cuda/reduce.cu.  Every predicate of the reference is
// kept; only the evaluation ORDER of side-effect-free predicates is changed (cheap tests first).
#pragma once

#include "ef_math.cuh"

namespace ef
{

// dense or pitched 3-plane SoA map: component c of pixel (x, y) lives at row y + c*rows
struct Map3
{
    const float * p;
    int pitch; // in floats
    int rows;
    __device__ __forceinline__ const float * row(int c, int y) const { return p + (size_t)(y + c * rows) * pitch; }
};

struct IcpParams
{
    Mat33 Rcurr;
    float3 tcurr;
    Mat33 Rprev_inv;
    float3 tprev;
    Intr intr;
    float dist_thresh, angle_thresh;
    int rows, cols;
};

struct RgbResParams
{
    Mat33 krkinv;
    float3 kt;
    float min_scale;
    float max_depth_delta;
    int rows, cols;
};

struct RgbStepParams
{
    float sigma;
    float fx, fy;           // level intrinsics
    float inv_fx, inv_fy;   // 1.0f / fx on the host (cudafuncs.cu:671)
    float cx, cy;
    float sobel_scale;
};

struct So3Params
{
    Mat33 image_basis, kinv, krlr;
    int rows, cols;
};

// 27 upper-triangle products + residual (row[6]^2) + inlier count: types.cuh:101-152 order
__device__ __forceinline__ void accumulate_se3(float * acc, const float * row)
{
    int k = 0;
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
#pragma unroll
        for(int j = i; j < 7; j++) acc[k++] += row[i] * row[j];
    }
    acc[27] += row[6] * row[6];
    acc[28] += 1.0f;
}

// 9 products + residual + inliers: types.cuh:154-181
__device__ __forceinline__ void accumulate_so3(float * acc, const float * row)
{
    int k = 0;
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
#pragma unroll
        for(int j = i; j < 4; j++) acc[k++] += row[i] * row[j];
    }
    acc[9] += row[3] * row[3];
    acc[10] += 1.0f;
}

// ICPReduction::search, first half (reduce.cu:285-298): transform the current vertex into the global and the
// previous-camera frame and project it.  Returns false when the projection leaves the image or lies behind
// the camera (:297).  A NaN vertex projects to pixel (0,0) (cvt.rni of NaN) and is rejected later by `dist`.
__device__ __forceinline__ bool icp_project(const IcpParams & P, const float3 & vcurr, float3 & vcurr_g, int & ux, int & uy)
{
    vcurr_g = P.Rcurr * vcurr + P.tcurr;
    const float3 vcurr_cp = P.Rprev_inv * (vcurr_g - P.tprev);
    ux = __float2int_rn(vcurr_cp.x * P.intr.fx / vcurr_cp.z + P.intr.cx); // :294
    uy = __float2int_rn(vcurr_cp.y * P.intr.fy / vcurr_cp.z + P.intr.cy);
    return !(ux < 0 || uy < 0 || ux >= P.cols || uy >= P.rows || vcurr_cp.z < 0); // :297
}

// ICPReduction::search, second half + getProducts (reduce.cu:310-347) once the model vertex / normal at the
// projected pixel are known.  Returns true and fills row[7] when the correspondence is accepted.
__device__ __forceinline__ bool icp_finish(const IcpParams & P, const float3 & vcurr_g, const float3 & ncurr, const float3 & vprev_g,
                                           const float3 & nprev_g, float * row)
{
    const float3 ncurr_g = P.Rcurr * ncurr;
    const float dist = norm3(vprev_g - vcurr_g);           // :317
    const float sine = norm3(cross3(ncurr_g, nprev_g));    // :318
    if(!(sine < P.angle_thresh && dist <= P.dist_thresh && !isnan(ncurr.x) && !isnan(nprev_g.x))) return false; // :324

    const float3 s_cp = P.Rprev_inv * (vcurr_g - P.tprev); // :341-343
    const float3 d_cp = P.Rprev_inv * (vprev_g - P.tprev);
    const float3 n_cp = P.Rprev_inv * nprev_g;
    const float3 c = cross3(s_cp, n_cp);
    row[0] = n_cp.x; row[1] = n_cp.y; row[2] = n_cp.z;
    row[3] = c.x; row[4] = c.y; row[5] = c.z;
    row[6] = dot3(n_cp, s_cp - d_cp);                      // :347
    return true;
}

// icp_finish without control flow: the row is always computed (garbage in, garbage out) and the verdict is returned,
// so that the caller can interleave several pixels and zero the rows of rejected ones with selects.  Same expressions,
// same bits as icp_finish for accepted pixels.
__device__ __forceinline__ bool icp_finish_select(const IcpParams & P, const float3 & vcurr_g, const float3 & ncurr, const float3 & vprev_g,
                                                  const float3 & nprev_g, float * row)
{
    const float3 ncurr_g = P.Rcurr * ncurr;
    const float dist = norm3(vprev_g - vcurr_g);           // :317
    const float sine = norm3(cross3(ncurr_g, nprev_g));    // :318
    const bool ok = sine < P.angle_thresh && dist <= P.dist_thresh && !isnan(ncurr.x) && !isnan(nprev_g.x); // :324
    const float3 s_cp = P.Rprev_inv * (vcurr_g - P.tprev); // :341-343
    const float3 d_cp = P.Rprev_inv * (vprev_g - P.tprev);
    const float3 n_cp = P.Rprev_inv * nprev_g;
    const float3 c = cross3(s_cp, n_cp);
    row[0] = n_cp.x; row[1] = n_cp.y; row[2] = n_cp.z;
    row[3] = c.x; row[4] = c.y; row[5] = c.z;
    row[6] = dot3(n_cp, s_cp - d_cp);                      // :347
    return ok;
}

// both halves with the gather in between (stand-alone operator kernel)
__device__ __forceinline__ bool icp_row(const IcpParams & P, const float3 & vcurr, const float3 & ncurr, const Map3 & vprev,
                                        const Map3 & nprev, float * row)
{
    float3 vcurr_g;
    int ux, uy;
    if(!icp_project(P, vcurr, vcurr_g, ux, uy)) return false;
    float3 vprev_g, nprev_g;
    vprev_g.x = __ldg(vprev.row(0, uy) + ux);
    vprev_g.y = __ldg(vprev.row(1, uy) + ux);
    vprev_g.z = __ldg(vprev.row(2, uy) + ux);
    nprev_g.x = __ldg(nprev.row(0, uy) + ux);
    nprev_g.y = __ldg(nprev.row(1, uy) + ux);
    nprev_g.z = __ldg(nprev.row(2, uy) + ux);
    return icp_finish(P, vcurr_g, ncurr, vprev_g, nprev_g, row);
}

// RGBResidual::getProducts, the warp of pixel (x, y) with depth d1 into the last image (reduce.cu:813-817).
// The reference binary evaluates  k.x*x + k.y*y + k.z  as  fma(k.x, x, k.y*y) + k.z  and the outer d1*(..) + kt as
// one FMA (its PTX); pinned with explicit FMAs because a 1-ulp difference flips the round-to-nearest pixel
// index at exact .5 ties.  Returns false when the warped pixel leaves the image.
__device__ __forceinline__ bool rgb_warp(const RgbResParams & P, int x, int y, float d1, int & u0, int & v0, float & transformed_d1)
{
    const float xf = (float)x, yf = (float)y;
    const float s2 = __fmaf_rn(P.krkinv.r2.x, xf, P.krkinv.r2.y * yf) + P.krkinv.r2.z;
    const float s0 = __fmaf_rn(P.krkinv.r0.x, xf, P.krkinv.r0.y * yf) + P.krkinv.r0.z;
    const float s1 = __fmaf_rn(P.krkinv.r1.x, xf, P.krkinv.r1.y * yf) + P.krkinv.r1.z;
    transformed_d1 = __fmaf_rn(d1, s2, P.kt.z);
    u0 = __float2int_rn(__fmaf_rn(d1, s0, P.kt.x) / transformed_d1);
    v0 = __float2int_rn(__fmaf_rn(d1, s1, P.kt.y) / transformed_d1);
    return (u0 >= 0 && v0 >= 0 && u0 < P.cols && v0 < P.rows); // :817
}

// reduce.cu:821
__device__ __forceinline__ bool rgb_accept(const RgbResParams & P, float transformed_d1, float d0, uint8_t last)
{
    return d0 > 0 && fabsf(transformed_d1 - d0) <= P.max_depth_delta && last != 0;
}

// the cheap per-pixel gates of RGBResidual::getProducts that need no gather (reduce.cu:779-783, :800-811)
__device__ __forceinline__ bool rgb_gate(const RgbResParams & P, int x, int y, int valx, int valy, float d1)
{
    const int border = 16; // :779
    if(!(y >= border && y < P.rows - border && x >= border && x < P.cols - border)) return false; // :781
    if(!(x < P.cols - 5 && y < P.rows - 1)) return false;                                          // :783
    const float mTwo = (valx * valx) + (valy * valy); // :802 (int arithmetic, then to float)
    return (mTwo >= P.min_scale) && !isnan(d1);       // :804, :811
}

// RGBResidual::getProducts (reduce.cu:768-842) for pixel (x, y).  On success returns true and
// (u0, v0) = matched pixel in the last image, diff = I_next(x,y) - I_last(u0,v0), d0 = lastDepth(u0,v0).
// Image pitches are in ELEMENTS.
__device__ __forceinline__ bool rgb_residual_px(const RgbResParams & P, int x, int y, int valx, int valy, float d1,
                                                const uint8_t * __restrict__ next_image, int img_pitch,
                                                const uint8_t * __restrict__ last_image, const float * __restrict__ last_depth,
                                                int depth_pitch, int & u0, int & v0, float & diff, float & d0)
{
    if(!rgb_gate(P, x, y, valx, valy, d1)) return false;

    // :787-793  4x4 neighbourhood [y-2, y+2) x [x-2, x+2) of the next image must be non-zero
    // (inside the 16-pixel border the max/min clamps of the reference are no-ops)
    bool valid = true;
#pragma unroll
    for(int u = -2; u < 2; u++)
    {
        const uint8_t * r = next_image + (size_t)(y + u) * img_pitch + x;
#pragma unroll
        for(int v = -2; v < 2; v++) valid = valid && (__ldg(r + v) > 0);
    }
    if(!valid) return false;

    float transformed_d1;
    if(!rgb_warp(P, x, y, d1, u0, v0, transformed_d1)) return false;

    d0 = __ldg(last_depth + (size_t)v0 * depth_pitch + u0);
    const uint8_t l = __ldg(last_image + (size_t)v0 * img_pitch + u0);
    if(!rgb_accept(P, transformed_d1, d0, l)) return false; // :821

    diff = static_cast<float>(__ldg(next_image + (size_t)y * img_pitch + x)) - static_cast<float>(l); // :827
    return true;
}

// RGBReduction::getProducts (reduce.cu:512-595) for one valid correspondence.  (X, Y, Z) is the
// point-cloud entry of the matched pixel (projectPointsKernel, cudafuncs.cu:656-658).
__device__ __forceinline__ void rgb_row(const RgbStepParams & P, float diff, float X, float Y, float Z, short gx, short gy,
                                        float * row)
{
    float w = P.sigma + fabsf(diff);                  // :523
    w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;        // :525
    if(P.sigma == -1) w = 1;                          // :528

    row[6] = -w * diff;                               // :533

    // :539 is a DOUBLE reciprocal rounded to float.  That is the correctly rounded float reciprocal: rounding first to 53 bits and then
    // to 24 cannot differ from rounding once when 53 >= 2 * 24 + 2 (checked over all 2^23 mantissas) -- and a float reciprocal
    // costs a fifth of the FP64 sequence on this part (FP64 issues at ~6.5 cycles per warp instruction)
    const float invz = __frcp_rn(Z);
    const float dI_dx_val = w * P.sobel_scale * gx;   // :540
    const float dI_dy_val = w * P.sobel_scale * gy;
    const float v0 = dI_dx_val * P.fx * invz;
    const float v1 = dI_dy_val * P.fy * invz;
    const float v2 = -(v0 * X + v1 * Y) * invz;

    row[0] = v0;
    row[1] = v1;
    row[2] = v2;
    row[3] = -Z * v1 + Y * v2;                        // :549-551
    row[4] = Z * v0 - X * v2;
    row[5] = -Y * v0 + X * v1;
}

// projectPointsKernel (cudafuncs.cu:641-659) for one pixel
__device__ __forceinline__ float3 project_point(int x, int y, float z, float inv_fx, float inv_fy, float cx, float cy)
{
    float3 p;
    p.x = (float)((x - cx) * z * inv_fx);
    p.y = (float)((y - cy) * z * inv_fy);
    p.z = z;
    return p;
}

// SO3Reduction::getGradient (reduce.cu:954-970); pitch in elements
__device__ __forceinline__ float2 so3_gradient(const uint8_t * __restrict__ img, int pitch, int x, int y)
{
    float2 g;
    const float actu = static_cast<float>(__ldg(img + (size_t)y * pitch + x));
    float back = static_cast<float>(__ldg(img + (size_t)y * pitch + x - 1));
    float fore = static_cast<float>(__ldg(img + (size_t)y * pitch + x + 1));
    g.x = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    back = static_cast<float>(__ldg(img + (size_t)(y - 1) * pitch + x));
    fore = static_cast<float>(__ldg(img + (size_t)(y + 1) * pitch + x));
    g.y = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    return g;
}

// SO3Reduction::getProducts (reduce.cu:972-1055) for pixel (x, y)
__device__ __forceinline__ bool so3_row(const So3Params & P, int x, int y, const uint8_t * __restrict__ last_image,
                                        const uint8_t * __restrict__ next_image, int pitch, float * row)
{
    const float3 unwarped = {(float)x, (float)y, 1.0f};                 // :980
    const float3 warped = P.image_basis * unwarped;                     // :982
    const int wx = __float2int_rn(warped.x / warped.z);                 // :984-985
    const int wy = __float2int_rn(warped.y / warped.z);

    if(!(wx >= 1 && wx < P.cols - 1 && wy >= 1 && wy < P.rows - 1 && x >= 1 && x < P.cols - 1 && y >= 1 && y < P.rows - 1))
        return false;                                                   // :987-997

    const float2 gradNext = so3_gradient(next_image, pitch, wx, wy);
    const float2 gradLast = so3_gradient(last_image, pitch, x, y);
    const float gx = (gradNext.x + gradLast.x) / 2.0f;                  // :1007-1008
    const float gy = (gradNext.y + gradLast.y) / 2.0f;

    const float3 point = P.kinv * unwarped;
    const float z2 = point.z * point.z;

    const float a = P.krlr.r0.x, b = P.krlr.r0.y, c = P.krlr.r0.z;
    const float d = P.krlr.r1.x, e = P.krlr.r1.y, f = P.krlr.r1.z;
    const float g = P.krlr.r2.x, h = P.krlr.r2.y, i = P.krlr.r2.z;

    // :1027-1029
    const float3 leftProduct = {((point.z * (d * gy + a * gx)) - (gy * g * y) - (gx * g * x)) / z2,
                                ((point.z * (e * gy + b * gx)) - (gy * h * y) - (gx * h * x)) / z2,
                                ((point.z * (f * gy + c * gx)) - (gy * i * y) - (gx * i * x)) / z2};
    const float3 jacRow = cross3(leftProduct, point);

    row[0] = jacRow.x;
    row[1] = jacRow.y;
    row[2] = jacRow.z;
    row[3] = -(static_cast<float>(__ldg(next_image + (size_t)wy * pitch + wx)) -
               static_cast<float>(__ldg(last_image + (size_t)y * pitch + x))); // :1036
    return true;
}

} // namespace ef
