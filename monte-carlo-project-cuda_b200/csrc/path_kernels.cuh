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
// (inc/trajectories.cuh:304-305): every store touches its own 32-byte sector.
//
// Here the STEPS of a path are spread over the lanes of a warp, which the stateless generator
// allows (normal i of path p is a pure function of (seed, p, i)): lane j draws the SPL (4 or 8)
// consecutive normals of steps [SPL*j, SPL*j + SPL) from its own Philox block(s), forms the
// in-lane prefix of the log2 increments, a 5-stage shuffle scan over the lane totals supplies
// each lane's starting log-price, and the warp then stores 32*SPL consecutive floats of ONE row:
// fully coalesced STG.128 with no shared-memory staging, no transposition and no barrier.
// Rows longer than 32*SPL steps take several passes with the running log-price carried in a
// register.  The summation order is a function of the step index only, so a row does not depend
// on which launch, slab or GPU produced it.
// ------------------------------------------------------------------------------------------
struct PathParams {
    float l0, dz, v, lB;
    int n_steps;
    uint32_t pad;
    uint64_t first_path;
    uint64_t n_paths;   // paths in this launch
    PhiloxKeys keys;
};

constexpr int kPathWarps = 8;     // warps (= rows in flight) per CTA
constexpr int kPathsPerWarp = 8;  // rows a warp walks one after the other

// exclusive prefix over the lanes of a warp (lane 0 gets 0) and the warp total
template <typename T>
__device__ __forceinline__ T warp_exclusive_scan(T x, int lane, T &total)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T y = __shfl_up_sync(kFullMask, x, off);
        if (lane >= off) x = x + y;
    }
    total = __shfl_sync(kFullMask, x, 31);
    const T up = __shfl_up_sync(kFullMask, x, 1);
    return lane ? up : T(0);
}

// SPL = steps per lane (4: one Philox block, 8: two).  VEC4: rows are 16-byte aligned
// (n_steps % 4 == 0 and aligned base pointers).  logs (nullable): log2 of every stored price,
// the exact FP32 state nested_kernel restarts its inner paths from.
template <int SPL, bool VEC4, bool COUNTS>
__global__ void __launch_bounds__(kPathWarps * 32)
trajectory_kernel(const __grid_constant__ PathParams prm, float *__restrict__ prices, int *__restrict__ counts,
                  float *__restrict__ logs)
{
    constexpr int kBlocks = SPL / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_steps = prm.n_steps;
    const uint64_t row0 = ((uint64_t)blockIdx.x * kPathWarps + warp) * kPathsPerWarp;

#pragma unroll 1
    for (int r = 0; r < kPathsPerWarp; ++r) {
        const uint64_t row = row0 + r;
        if (row >= prm.n_paths) return;
        const uint64_t p = prm.first_path + row;
        const uint32_t p_lo = (uint32_t)p, p_hi = (uint32_t)(p >> 32);
        const uint64_t row_off = row * (uint64_t)n_steps;
        float carry_l = prm.l0;
        int carry_c = 0;

#pragma unroll 1
        for (int step0 = 0; step0 < n_steps; step0 += 32 * SPL) {
            const int my_step = step0 + SPL * lane;
            float a[SPL];  // in-lane inclusive prefix of the log2 increments
            if (my_step < n_steps) {
                float z[SPL];
#pragma unroll
                for (int b = 0; b < kBlocks; ++b)
                    normals4(philox4x32_10((uint32_t)(my_step >> 2) + b, 0u, p_lo, p_hi, prm.keys), prm.dz, z + 4 * b);
                a[0] = prm.v * z[0];
#pragma unroll
                for (int j = 1; j < SPL; ++j) a[j] = fmaf(prm.v, z[j], a[j - 1]);
            } else {
#pragma unroll
                for (int j = 0; j < SPL; ++j) a[j] = 0.0f;
            }
            float total;
            const float base = carry_l + warp_exclusive_scan(a[SPL - 1], lane, total);
            carry_l = carry_l + total;

            float s[SPL];
            int c[SPL];
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                a[j] = base + a[j];
                s[j] = mufu_ex2(a[j]);
            }
            if (COUNTS) {
                int run = 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    run += (a[j] < prm.lB && my_step + j < n_steps) ? 1 : 0;
                    c[j] = run;
                }
                int ctotal;
                const int cbase = carry_c + warp_exclusive_scan(run, lane, ctotal);
                carry_c += ctotal;
#pragma unroll
                for (int j = 0; j < SPL; ++j) c[j] += cbase;
            }

            if (my_step < n_steps) {
                const uint64_t off = row_off + (uint64_t)my_step;
                if (VEC4) {  // n_steps % 4 == 0: whole float4s are in range
#pragma unroll
                    for (int b = 0; b < kBlocks; ++b) {
                        if (my_step + 4 * b < n_steps) {
                            __stcs(reinterpret_cast<float4 *>(prices + off) + b,
                                   make_float4(s[4 * b], s[4 * b + 1], s[4 * b + 2], s[4 * b + 3]));
                            if (COUNTS)
                                __stcs(reinterpret_cast<int4 *>(counts + off) + b,
                                       make_int4(c[4 * b], c[4 * b + 1], c[4 * b + 2], c[4 * b + 3]));
                            if (logs)
                                __stcs(reinterpret_cast<float4 *>(logs + off) + b,
                                       make_float4(a[4 * b], a[4 * b + 1], a[4 * b + 2], a[4 * b + 3]));
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        if (my_step + j < n_steps) {
                            __stcs(prices + off + j, s[j]);
                            if (COUNTS) __stcs(counts + off + j, c[j]);
                            if (logs) __stcs(logs + off + j, a[j]);
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Nested Monte Carlo: one CTA owns one outer trajectory p and every inner path hanging off it.
// The outer walk is trajectory_kernel's (it also leaves the exact FP32 log2-price and barrier
// count of every point in a workspace); here the CTA reads its row of point states (400 bytes)
// and runs the inner conditional-expectation paths entirely in registers; the per-point
// (sum, sumsq) fold through the same fixed tree as everything else -- no atomics (the
// reference: inc/nmc.cuh:100-104, 378-381), nothing but F[p,k] is written.
// Inner path j of point q = p*n_steps + k draws from stream (seed_inner, q*n_inner + j) and
// RESTARTS from the point's state (the reference carries state over, inc/nmc.cuh:51-53).
// ------------------------------------------------------------------------------------------
struct NestedParams {
    float dz, v, lB, K;
    int P1, P2, n_steps, n_inner;
    int discount_mode;     // MCB_DISCOUNT_*
    float r, T, dt;
    uint64_t first_outer;
    PhiloxKeys keys_inner;
};

__global__ void __launch_bounds__(kSlots)
nested_kernel(const __grid_constant__ NestedParams prm, const float *__restrict__ logs,
              const int *__restrict__ counts, float *__restrict__ F)
{
    __shared__ float scratch[2 * kWarps];
    const uint64_t p = prm.first_outer + blockIdx.x;
    const int n_steps = prm.n_steps;
    const uint64_t row = (uint64_t)blockIdx.x * (uint64_t)n_steps;

#pragma unroll 1
    for (int k = 0; k < n_steps; ++k) {
        const float lo = __ldg(logs + row + k);
        const int co = __ldg(counts + row + k);
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

}  // namespace mcb
