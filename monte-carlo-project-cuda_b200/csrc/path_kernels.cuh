// path_kernels.cuh -- full-trajectory storage and nested Monte Carlo.  sm_100a only.
//
// Replaces (reference file:line, relative to the reference repo)
//   simulate_outer_trajectories                         inc/trajectories.cuh:273-351
//   simulateOptionPriceMultipleBlockGPU (overload B)    inc/testing.cuh:46-73
//   compute_nmc_one_block_per_point / _with_outter / compute_nmc_optimal   inc/nmc.cuh:12-386
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "block_reduce.cuh"
#include "philox.cuh"
#include "pricing_kernels.cuh"

namespace mcb {

// ------------------------------------------------------------------------------------------
// Trajectory store, path-major prices[(p - first_path) * n_steps + i] = S(t_{i+1}).
// The reference stores one float per thread per step with a warp stride of n_steps*4 bytes
// (inc/trajectories.cuh:304-305): every store touches its own 32-byte sector.  Here a warp
// owns 32 consecutive paths, walks them 32 steps at a time (8 Philox blocks per lane),
// stages the 32x32 tile in shared memory with conflict-free 128-bit stores, and writes it
// out transposed: each quarter-warp stores 128 contiguous bytes of one path's row
// (STG.128), so every sector written is full.
// ------------------------------------------------------------------------------------------
struct PathParams {
    float l0, dz, v, lB;
    int n_steps;
    uint32_t pad;
    uint64_t first_path;
    uint64_t n_paths;   // paths in this launch
    PhiloxKeys keys;
};

constexpr int kTileSteps = 32;
constexpr int kTileStride = 36;   // floats per staged row: 144 B keeps STS.128/LDS.128 conflict-free
constexpr int kPathWarps = 4;     // warps per CTA: 18 KB (prices) / 36 KB (+counts) of static smem

template <bool VEC4, bool COUNTS>
__global__ void __launch_bounds__(kPathWarps * 32)
trajectory_kernel(const __grid_constant__ PathParams prm, float *__restrict__ prices, int *__restrict__ counts)
{
    __shared__ __align__(16) float tile_p[kPathWarps][32][kTileStride];
    __shared__ __align__(16) int tile_c[COUNTS ? kPathWarps : 1][COUNTS ? 32 : 1][kTileStride];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t local0 = ((uint64_t)blockIdx.x * kPathWarps + warp) * 32;  // first path of this warp (launch-local)
    if (local0 >= prm.n_paths) return;
    const uint64_t p = prm.first_path + local0 + lane;
    const uint32_t p_lo = (uint32_t)p, p_hi = (uint32_t)(p >> 32);
    const int n_steps = prm.n_steps;

    float l = prm.l0;
    int count = 0;
    float(*tp)[kTileStride] = tile_p[warp];
    int(*tc)[kTileStride] = tile_c[COUNTS ? warp : 0];

    for (int step0 = 0; step0 < n_steps; step0 += kTileSteps) {
        const int steps_here = min(kTileSteps, n_steps - step0);
        const int nblk = (steps_here + 3) >> 2;
#pragma unroll 2
        for (int b = 0; b < nblk; ++b) {
            float z[4];
            normals4(philox4x32_10((uint32_t)((step0 >> 2) + b), 0u, p_lo, p_hi, prm.keys), prm.dz, z);
            float4 s;
            int4 c;
            l = fmaf(prm.v, z[0], l); s.x = mufu_ex2(l); if (COUNTS) { count += (l < prm.lB) ? 1 : 0; c.x = count; }
            l = fmaf(prm.v, z[1], l); s.y = mufu_ex2(l); if (COUNTS) { count += (l < prm.lB) ? 1 : 0; c.y = count; }
            l = fmaf(prm.v, z[2], l); s.z = mufu_ex2(l); if (COUNTS) { count += (l < prm.lB) ? 1 : 0; c.z = count; }
            l = fmaf(prm.v, z[3], l); s.w = mufu_ex2(l); if (COUNTS) { count += (l < prm.lB) ? 1 : 0; c.w = count; }
            *reinterpret_cast<float4 *>(&tp[lane][4 * b]) = s;
            if (COUNTS) *reinterpret_cast<int4 *>(&tc[lane][4 * b]) = c;
        }
        __syncwarp();
        if (VEC4) {
            // n_steps % 4 == 0: rows are 16-byte aligned and whole float4s are in range.
            const int c4 = lane & 7;
            const int step = step0 + 4 * c4;
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int row = rr * 4 + (lane >> 3);
                if (local0 + row < prm.n_paths && step < n_steps) {
                    const uint64_t off = (local0 + row) * (uint64_t)n_steps + (uint64_t)step;
                    __stcs(reinterpret_cast<float4 *>(prices + off), *reinterpret_cast<const float4 *>(&tp[row][4 * c4]));
                    if (COUNTS)
                        __stcs(reinterpret_cast<int4 *>(counts + off), *reinterpret_cast<const int4 *>(&tc[row][4 * c4]));
                }
            }
        } else {
            const int step = step0 + lane;
            for (int row = 0; row < 32; ++row) {
                if (local0 + row < prm.n_paths && step < n_steps) {
                    const uint64_t off = (local0 + row) * (uint64_t)n_steps + (uint64_t)step;
                    __stcs(prices + off, tp[row][lane]);
                    if (COUNTS) __stcs(counts + off, tc[row][lane]);
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Nested Monte Carlo: one CTA owns one outer trajectory p and every inner path hanging off
// it.  The outer state (log2 S, I) is advanced in registers (redundantly per thread: 100
// steps against ~8e4 inner steps per thread), so the inner conditional-expectation paths
// never read global memory; the per-point (sum, sumsq) fold through the same fixed tree as
// everything else -- no atomics (the reference: inc/nmc.cuh:100-104, 378-381).
// Inner path j of point q = p*n_steps + k draws from stream (seed_inner, q*n_inner + j) and
// RESTARTS from the point's state (the reference carries state over, inc/nmc.cuh:51-53).
// ------------------------------------------------------------------------------------------
struct NestedParams {
    float l0, dz, v, lB, K;
    int P1, P2, n_steps, n_inner;
    int discount_mode;     // MCB_DISCOUNT_*
    float r, T, dt;
    uint32_t pad;
    uint64_t first_outer;
    PhiloxKeys keys_outer;
    PhiloxKeys keys_inner;
};

__global__ void __launch_bounds__(kSlots)
nested_kernel(const __grid_constant__ NestedParams prm, float *__restrict__ F, float *__restrict__ prices,
              int *__restrict__ counts)
{
    __shared__ float scratch[2 * kWarps];
    const uint64_t p = prm.first_outer + blockIdx.x;
    const uint32_t p_lo = (uint32_t)p, p_hi = (uint32_t)(p >> 32);
    const int n_steps = prm.n_steps;
    const uint64_t row = (uint64_t)blockIdx.x * (uint64_t)n_steps;

    float lo = prm.l0;
    int co = 0;
    for (int k4 = 0; k4 < n_steps; k4 += 4) {
        float zo[4];
        normals4(philox4x32_10((uint32_t)(k4 >> 2), 0u, p_lo, p_hi, prm.keys_outer), prm.dz, zo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k4 + j;
            if (k < n_steps) {
                lo = fmaf(prm.v, zo[j], lo);
                co += (lo < prm.lB) ? 1 : 0;
                if (threadIdx.x == 0) {
                    if (prices) prices[row + k] = mufu_ex2(lo);
                    if (counts) counts[row + k] = co;
                }
                const int remaining = n_steps - (k + 1);
                float sum = 0.0f, sq = 0.0f;
                if (co <= prm.P2) {
                    const uint64_t q = (p * (uint64_t)n_steps + (uint64_t)k) * (uint64_t)prm.n_inner;
#pragma unroll 1
                    for (int jj = threadIdx.x; jj < prm.n_inner; jj += kSlots) {
                        const uint64_t sub = q + (uint64_t)jj;
                        float l = lo;
                        int c = co;
                        walk_path(l, c, (uint32_t)sub, (uint32_t)(sub >> 32), remaining, prm.dz, prm.v, prm.lB,
                                  prm.keys_inner);
                        const float pay = (c >= prm.P1 && c <= prm.P2) ? fmaxf(mufu_ex2(l) - prm.K, 0.0f) : 0.0f;
                        sum = sum + pay;
                        sq = fmaf(pay, pay, sq);
                    }
                }
                block_fold2(sum, sq, scratch);
                if (threadIdx.x == 0) {
                    const double tau = prm.discount_mode == 1 ? (double)prm.T - (double)(k + 1) * (double)prm.dt
                                                              : (double)prm.T;
                    const double scale = exp(-(double)prm.r * tau) / (double)prm.n_inner;
                    F[row + k] = (float)(scale * (double)sum);
                }
                __syncthreads();  // scratch is reused by the next point
            }
        }
    }
}

}  // namespace mcb
