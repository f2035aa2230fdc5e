/*
 * oracle/ref_shim.h -- TEST INFRASTRUCTURE (never linked into the product library).
 *
 * Force-included (nvcc -include) in front of the UNMODIFIED reference sources
 *   /root/reference/elasticfusionpublic/Core/src/Cuda/{reduce.cu,cudafuncs.cu}
 * so that they compile with CUDA 12.9 for sm_100a where they lie (no copy, no patch):
 *
 *  1. reduce.cu:94-128,191-204,684-685 call the pre-Volta `__shfl_down(x, offset)`
 *     which no longer exists for sm_70+.  We supply device overloads with the same
 *     signature that forward to `__shfl_down_sync(0xffffffff, ...)` -- identical data
 *     movement for the full, converged warps the reference always uses.
 *  2. cudafuncs.cu:548-577,681-711 use legacy texture *references*
 *     (`texture<uchar4,2> inTex; cudaBindTextureToArray; tex2D(inTex,x,y)`), removed in
 *     CUDA 12.  We supply a tiny `texture<>` stand-in that carries a
 *     cudaTextureObject_t in a __device__ variable; bind == create object + copy to
 *     symbol.  The arithmetic at cudafuncs.cu:560 is untouched.
 */
#ifndef EF_ORACLE_REF_SHIM_H_
#define EF_ORACLE_REF_SHIM_H_

#include <cuda_runtime.h>
#include <cstring>

#if defined(__CUDA_ARCH__)
/* device pass only: in the host pass reduce.cu:56-80 supplies its own (never executed)
 * fallback definitions because __CUDA_ARCH__ is undefined there. */
static __device__ __forceinline__ float __shfl_down(float v, int offset)
{
    return __shfl_down_sync(0xffffffffu, v, offset);
}
static __device__ __forceinline__ int __shfl_down(int v, int offset)
{
    return __shfl_down_sync(0xffffffffu, v, offset);
}
#endif

template<class T, int dim, cudaTextureReadMode mode>
struct ef_ref_texture
{
    cudaTextureObject_t obj;
};

/* `texture<uchar4, 2, cudaReadModeElementType> inTex;` at file scope becomes a
 * __device__ variable holding a texture object. */
#define texture __device__ ef_ref_texture

template<class T, int dim, cudaTextureReadMode mode>
static __device__ __forceinline__ T tex2D(const ef_ref_texture<T, dim, mode> & t, int x, int y)
{
    return tex2D<T>(t.obj, (float)x, (float)y);
}

template<class T, int dim, cudaTextureReadMode mode>
static inline cudaError_t cudaBindTextureToArray(const ef_ref_texture<T, dim, mode> & sym, cudaArray * arr)
{
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;

    /* legacy texture references defaulted to point filtering, clamp addressing,
     * unnormalised coordinates, element read mode */
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeClamp;
    td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = mode;
    td.normalizedCoords = 0;

    ef_ref_texture<T, dim, mode> host;
    cudaError_t err = cudaCreateTextureObject(&host.obj, &res, &td, NULL);
    if(err != cudaSuccess) return err;
    return cudaMemcpyToSymbol(sym, &host, sizeof(host));
}

template<class T, int dim, cudaTextureReadMode mode>
static inline cudaError_t cudaUnbindTexture(const ef_ref_texture<T, dim, mode> & sym)
{
    ef_ref_texture<T, dim, mode> host;
    cudaError_t err = cudaMemcpyFromSymbol(&host, sym, sizeof(host));
    if(err != cudaSuccess) return err;
    return cudaDestroyTextureObject(host.obj);
}

#endif /* EF_ORACLE_REF_SHIM_H_ */
