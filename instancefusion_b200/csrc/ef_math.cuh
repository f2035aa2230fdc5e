// ef_math.cuh -- small device-side vector helpers and per-pixel math of the tracker.
//
// The expression SHAPES here deliberately follow the reference helpers
// (elasticfusionpublic/Core/src/Cuda/operators.cuh:55-91) so that, compiled with the reference's
// flags (--ftz=true --prec-div=false --prec-sqrt=false, default --fmad=true), nvcc emits the same
// FMA contractions / approximate ops and float results agree bit-for-bit with the reference kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ef
{

struct Mat33
{
    float3 r0, r1, r2; // rows
};

struct Intr
{
    float fx, fy, cx, cy;
};

__device__ __forceinline__ float3 operator-(const float3 & a, const float3 & b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator+(const float3 & a, const float3 & b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
// a.x*b.x + a.y*b.y + a.z*b.z.  nvcc contracts this as  t = a.y*b.y (FMUL); t = fma(a.x, b.x, t); t = fma(a.z, b.z, t)
// in every reference kernel (checked in its PTX: icpKernel, tranformMapsKernel), but which product it leaves
// un-fused depends on use counts in the surrounding code -- so the order is pinned here with explicit FMAs.
__device__ __forceinline__ float dot3(const float3 & a, const float3 & b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, a.y * b.y)); }
__device__ __forceinline__ float3 cross3(const float3 & a, const float3 & b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm3(const float3 & a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 normalized3(const float3 & a)
{
    const float rn = rsqrtf(dot3(a, a));
    return make_float3(a.x * rn, a.y * rn, a.z * rn);
}
__device__ __forceinline__ float3 operator*(const Mat33 & m, const float3 & a) { return make_float3(dot3(m.r0, a), dot3(m.r1, a), dot3(m.r2, a)); }

__device__ __forceinline__ float qnan() { return __int_as_float(0x7fffffff); }

// streaming (read-once) and read-only loads
template<class T> __device__ __forceinline__ T ldg(const T * p) { return __ldg(p); }

} // namespace ef
