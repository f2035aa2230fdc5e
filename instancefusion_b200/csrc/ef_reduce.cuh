// ef_reduce.cuh -- deterministic fixed-order sum reductions for the 29-float (SE3) and 11-float (SO3)
// normal-equation accumulators.
//
// The reference (reduce.cu:90-255) does 29 x 5 __shfl_down per warp, twice per block, then a second
// <<<1,1024>>> kernel, and its float result depends on the (threads, blocks) launch shape.
// Here:
//   warp   : "transpose-reduce" butterfly -- at offset 16 every lane hands HALF of its values to its
//            partner and keeps the other half, at offset 8 a quarter, ...  After 5 steps lane L holds
//            the warp total of value index L.  31 shuffles instead of 145, and a fixed tree.
//   block  : lane L of every warp stores its total to smem[warp][L]; warp 0 adds the warps in index order.
//   grid   : every block writes a 32-float partial row; the LAST block to arrive (atomic ticket +
//            __threadfence) adds the rows in block-index order.  No second launch.
// The summation tree is a pure function of (blockDim, gridDim, pixel->thread map): deterministic.
#pragma once

#include <cuda_runtime.h>

namespace ef
{

constexpr unsigned kFullMask = 0xffffffffu;

// v[0..31]: per-lane values (pad unused slots with 0).  On return, v[0] of lane L is the warp-wide sum of
// slot L.
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32])
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for(int half = 16; half >= 1; half >>= 1)
    {
        const bool up = (lane & half) != 0;
#pragma unroll
        for(int i = 0; i < half; i++)
        {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFullMask, send, half);
        }
    }
    return v[0];
}

// 16-slot variant for the SO3 accumulator (11 used): lanes L and L+16 both end with slot (L & 15) of
// their half-warp pair sum; one more xor-16 add makes it the full-warp total.
__device__ __forceinline__ float warp_transpose_reduce16(float (&v)[16])
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for(int half = 8; half >= 1; half >>= 1)
    {
        const bool up = (lane & half) != 0;
#pragma unroll
        for(int i = 0; i < half; i++)
        {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFullMask, send, half);
        }
    }
    // lanes 0..15 hold slot (lane) summed over lanes with equal low-4 bits in their 16-lane half
    const float other = __shfl_xor_sync(kFullMask, v[0], 16);
    // fixed order: lower half first
    return (lane & 16) ? (other + v[0]) : (v[0] + other);
}

// Block-level: every thread passes the lane value returned by the warp reduce (slot = lane).  Returns
// the block total of slot `lane` in warp 0 (other warps return garbage).  smem: float[32][32].
template<int SLOTS>
__device__ __forceinline__ float block_reduce_slots(float lane_value, float * smem)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned nwarps = (blockDim.x + 31u) >> 5;
    if(lane < SLOTS) smem[warp * SLOTS + lane] = lane_value;
    __syncthreads();
    float total = 0.f;
    if(warp == 0 && lane < SLOTS)
    {
        for(unsigned w = 0; w < nwarps; w++) total += smem[w * SLOTS + lane];
    }
    return total;
}

// Grid-level last-block-done reduction of `SLOTS`-wide partial rows (row stride 32 floats).
// Called by ALL threads of every block; `block_total` is meaningful in warp 0, lanes < SLOTS.
// partials: gridDim.x * 32 floats; ticket: one unsigned, must be 0 on entry and is reset to 0.
// Returns true in the threads of the last block, after which out[0..SLOTS) is final (written by
// this call) -- the caller may run an epilogue there.
template<int SLOTS>
__device__ __forceinline__ bool grid_reduce_last_block(float block_total, float * __restrict__ partials, unsigned * ticket,
                                                       float * __restrict__ out, float * smem)
{
    __shared__ bool is_last;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    if(warp == 0 && lane < SLOTS) partials[blockIdx.x * 32 + lane] = block_total;
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if(!is_last) return false;
    __threadfence();

    // fixed order: warp w adds rows w, w+nwarps, ... ; then warp 0 adds the warps in index order
    const unsigned nwarps = (blockDim.x + 31u) >> 5;
    // (kTailBatch rows of a lane are requested together and then added in the fixed order: the tail of the kernel is a chain
    //  of L2 round trips, one per batch instead of one per row)
    constexpr unsigned kTailBatch = 16;
    float s = 0.f;
    if(lane < SLOTS)
        for(unsigned b0 = warp; b0 < gridDim.x; b0 += kTailBatch * nwarps)
        {
            float v[kTailBatch];
#pragma unroll
            for(unsigned k = 0; k < kTailBatch; k++)
            {
                const unsigned b = b0 + k * nwarps;
                v[k] = b < gridDim.x ? __ldcg(partials + b * 32 + lane) : 0.f;
            }
#pragma unroll
            for(unsigned k = 0; k < kTailBatch; k++)
                if(b0 + k * nwarps < gridDim.x) s += v[k];
        }
    __syncthreads(); // smem reuse
    if(lane < SLOTS) smem[warp * SLOTS + lane] = s;
    __syncthreads();
    if(warp == 0 && lane < SLOTS)
    {
        float total = 0.f;
        for(unsigned w = 0; w < nwarps; w++) total += smem[w * SLOTS + lane];
        out[lane] = total;
    }
    if(threadIdx.x == 0) *ticket = 0u;
    __syncthreads();
    return true;
}

} // namespace ef
