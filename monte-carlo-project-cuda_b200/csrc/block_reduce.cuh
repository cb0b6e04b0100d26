// block_reduce.cuh -- deterministic warp-shuffle + shared-memory reduction of (sum, sumsq).
//
// Replaces reduce3/4/5/6 (inc/reduce.cuh:9-227) and the copy of that tree inlined in every
// reference kernel (inc/trajectories.cuh:77-111 etc.: five block barriers, float sum only,
// one float atomicAdd per block).  Here: lanes fold 16,8,4,2,1 by shuffle, the 8 warp
// results fold 4,2,1 -- one barrier, no atomics, and a sum of squares alongside the sum so
// the engine can report a standard error.  The operation order is a function of the slot
// index only, so the result does not depend on launch geometry; oracle/mc_oracle.c
// (tree256_f32 / tree256_f64) restates it bit for bit.
#pragma once
#include <cuda_runtime.h>

namespace mcb {

constexpr int kSlots = 256;       // MCB_SLOTS
constexpr int kWarps = kSlots / 32;
constexpr unsigned kFullMask = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T warp_fold(T v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_down_sync(kFullMask, v, off);
    return v;  // lane 0 holds the warp total
}

// Reduce two values over the 256 threads of a CTA.  Thread 0 returns the totals; other
// threads return unspecified values.  `scratch` must hold 2*kWarps elements of T and may be
// reused after the call returns on thread 0 only once the CTA has synchronised again.
template <typename T>
__device__ __forceinline__ void block_fold2(T &a, T &b, T *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_fold(a);
    b = warp_fold(b);
    if (lane == 0) {
        scratch[warp] = a;
        scratch[kWarps + warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
        T x = lane < kWarps ? scratch[lane] : T(0);
        T y = lane < kWarps ? scratch[kWarps + lane] : T(0);
#pragma unroll
        for (int off = kWarps / 2; off > 0; off >>= 1) {
            x = x + __shfl_down_sync(kFullMask, x, off);
            y = y + __shfl_down_sync(kFullMask, y, off);
        }
        a = x;
        b = y;
    }
}

}  // namespace mcb
