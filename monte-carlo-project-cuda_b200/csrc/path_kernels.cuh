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
// Here the STEPS of a path are spread over LPR lanes of a warp (32, or 16 with two rows side by
// side), which the stateless generator allows (normal i of path p is a pure function of
// (seed, p, i)): lane j draws the SPL consecutive normals of steps [SPL*j, SPL*j + SPL) from its
// own Philox block(s), forms the in-lane prefix of the log2 increments, a log2(LPR)-stage
// shuffle scan over the lane totals supplies each lane's starting log-price, and the lanes then
// store LPR*SPL consecutive floats of ONE row: every lane writes SPL*4 contiguous bytes
// (whole 32-byte sectors) with STG.128, no shared-memory staging, no transposition, no barrier.
// Rows longer than LPR*SPL steps take several passes with the running log-price carried in a
// register.  The summation order is a function of the step index only, so a row does not depend
// on which launch, slab or GPU produced it.
// ------------------------------------------------------------------------------------------
struct PathParams {
    float l0, sc, dr, lB;   // sc, dr: see WalkParams
    int n_steps;
    uint32_t pad;
    uint64_t first_path;
    uint64_t n_paths;   // paths in this launch
    PhiloxKeys keys;
};

constexpr int kPathWarps = 8;     // warps (= rows in flight) per CTA
constexpr int kPathsPerWarp = 8;  // rows a warp walks one after the other

// One stage of an inclusive scan over groups of LPR lanes: x += (value of lane - off) when
// that lane is in the same group.  shfl.sync.up hands back the "source lane in range" predicate,
// so a stage is SHFL + one predicated add (the C++ intrinsic costs SHFL + FADD + FSEL).
template <int LPR>
__device__ __forceinline__ float scan_stage(float x, int off)
{
    asm("{ .reg .f32 t; .reg .pred p;\n\t"
        "shfl.sync.up.b32 t|p, %0, %1, %2, 0xffffffff;\n\t"
        "@p add.f32 %0, %0, t; }"
        : "+f"(x) : "r"(off), "n"((32 - LPR) << 8));
    return x;
}
template <int LPR>
__device__ __forceinline__ int scan_stage(int x, int off)
{
    asm("{ .reg .b32 t; .reg .pred p;\n\t"
        "shfl.sync.up.b32 t|p, %0, %1, %2, 0xffffffff;\n\t"
        "@p add.s32 %0, %0, t; }"
        : "+r"(x) : "r"(off), "n"((32 - LPR) << 8));
    return x;
}
// value of the lane below (0 for the first lane of a group)
template <int LPR>
__device__ __forceinline__ float shift_up_one(float x)
{
    float r;
    asm("{ .reg .pred p;\n\t"
        "shfl.sync.up.b32 %0|p, %1, 1, %2, 0xffffffff;\n\t"
        "@!p mov.f32 %0, 0f00000000; }"
        : "=f"(r) : "f"(x), "n"((32 - LPR) << 8));
    return r;
}
template <int LPR>
__device__ __forceinline__ int shift_up_one(int x)
{
    int r;
    asm("{ .reg .pred p;\n\t"
        "shfl.sync.up.b32 %0|p, %1, 1, %2, 0xffffffff;\n\t"
        "@!p mov.b32 %0, 0; }"
        : "=r"(r) : "r"(x), "n"((32 - LPR) << 8));
    return r;
}
// exclusive prefix over each group of LPR lanes (first lane gets 0): 1 + log2(LPR) SHFL
template <int LPR, typename T>
__device__ __forceinline__ T group_exclusive_scan(T x)
{
    x = shift_up_one<LPR>(x);
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) x = scan_stage<LPR>(x, off);
    return x;
}

// ---- TMA bulk store helpers (cp.async.bulk shared::cta -> global, bulk_group completion) ----
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bulk_store(void *gmem, const void *smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gmem), "r"(smem_addr(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

enum { kStoreScalar = 0,   // any n_steps / alignment: 4-byte streaming stores
       kStoreVec4 = 1 };   // rows 16-byte aligned: STG.128 straight from registers

// One pass of one row: lane `ln` of the row's LPR lanes draws the SPL normals of steps
// [my_step, my_step + SPL), turns them into log2 increments, prefixes them in-lane and across the
// lanes (scan), and returns the log2 price BEFORE its first step in `base`; a[j] is the in-lane
// inclusive prefix, so step my_step + j ends at log2 price base + a[j].  carry_l (the log2 price
// at the start of the pass) is advanced to the end of the pass.  Both trajectory kernels go
// through this function, so a row's bits do not depend on which of them produced it.
template <int SPL>
struct PassWords {
    Words4 w[SPL / 4];
};

// integer half of a pass: the Philox blocks of steps [my_step, my_step + SPL) of path (p_lo, p_hi)
template <int SPL>
__device__ __forceinline__ PassWords<SPL> row_words(const PathParams &prm, uint32_t p_lo, uint32_t p_hi, int my_step)
{
    PassWords<SPL> out;
#pragma unroll
    for (int b = 0; b < SPL / 4; ++b) out.w[b] = philox4x32_10((uint32_t)(my_step >> 2) + b, 0u, p_lo, p_hi, prm.keys);
    return out;
}

// The first two Philox rounds of a trajectory block depend on far less than (lane, row): the
// counter is (block, 0, p_lo, p_hi), so
//   round 0:  M0 * block   -- a function of the LANE only (its blocks are the same for every row),
//             M1 * p_lo    -- a function of the ROW only (the same for the lane's SPL/4 blocks);
//   round 1:  M0 * (hi(M1 p_lo) ^ k0[0])            -- row only,
//             M1 * (hi(M0 block) ^ p_hi ^ k1[0])    -- lane only, as long as p_hi does not change.
// RowHoist keeps the lane-only products (three words per block) in registers for the whole kernel;
// a row then costs two multiplies for all its blocks and each block enters round 2 after two XORs:
// 16.5 instead of 20 IMAD.WIDE per block, the same words bit for bit.
template <int SPL>
struct RowHoist {
    uint32_t a[SPL / 4];   // hi(M1 * c2') ^ k0[1]
    uint32_t b[SPL / 4];   // lo(M1 * c2')
    uint32_t d[SPL / 4];   // lo(M0 * block) ^ k1[1]
    uint32_t p_hi;         // the high path word the products were formed with
};

template <int SPL>
__device__ __forceinline__ RowHoist<SPL> make_row_hoist(const PathParams &prm, uint32_t p_hi, int my_step)
{
    RowHoist<SPL> h;
    h.p_hi = p_hi;
#pragma unroll
    for (int j = 0; j < SPL / 4; ++j) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * ((uint32_t)(my_step >> 2) + (uint32_t)j);
        const uint32_t c2 = (uint32_t)(p0 >> 32) ^ p_hi ^ prm.keys.k1[0];
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        h.a[j] = (uint32_t)(p1 >> 32) ^ prm.keys.k0[1];
        h.b[j] = (uint32_t)p1;
        h.d[j] = (uint32_t)p0 ^ prm.keys.k1[1];
    }
    return h;
}

// == row_words(prm, p_lo, h.p_hi, my_step) for the my_step the hoist was made with
template <int SPL>
__device__ __forceinline__ PassWords<SPL> row_words_hoisted(const PathParams &prm, const RowHoist<SPL> &h, uint32_t p_lo)
{
    const uint64_t q1 = (uint64_t)kPhiloxM1 * p_lo;                                   // round 0, row part
    const uint64_t q0 = (uint64_t)kPhiloxM0 * ((uint32_t)(q1 >> 32) ^ prm.keys.k0[0]);  // round 1, row part
    PassWords<SPL> out;
#pragma unroll
    for (int j = 0; j < SPL / 4; ++j) {
        uint32_t c0 = h.a[j] ^ (uint32_t)q1, c1 = h.b[j], c2 = (uint32_t)(q0 >> 32) ^ h.d[j], c3 = (uint32_t)q0;
#pragma unroll
        for (int r = 2; r < 10; ++r) philox_round(c0, c1, c2, c3, prm.keys.k0[r], prm.keys.k1[r]);
        out.w[j] = Words4{c0, c1, c2, c3};
    }
    return out;
}

// floating-point half: Box-Muller -> log2 increments -> in-lane prefix -> scan over the row's lanes.
// PACK: the Box-Muller arithmetic in FP32x2 instructions (philox.cuh) -- fewer issue slots, same bits.
template <int SPL, int LPR, bool PACK = false>
__device__ __forceinline__ float row_finish(const PathParams &prm, const PassWords<SPL> &words, bool active,
                                            float &carry_l, float (&a)[SPL])
{
#pragma unroll
    for (int b = 0; b < SPL / 4; ++b) {
        if (PACK) increments4_packed(words.w[b], prm.sc, prm.dr, a + 4 * b);
        else increments4(words.w[b], prm.sc, prm.dr, a + 4 * b);
    }
#pragma unroll
    for (int j = 1; j < SPL; ++j) a[j] = a[j] + a[j - 1];
    // lanes past the end of the row computed garbage (branch-free, the warp stays converged for
    // the shuffles); they must not leak into the scan
    const float lane_total = active ? a[SPL - 1] : 0.0f;
    const float base = carry_l + group_exclusive_scan<LPR>(lane_total);
    carry_l = __shfl_sync(kFullMask, base + lane_total, LPR - 1, LPR);
    return base;
}

template <int SPL, int LPR>
__device__ __forceinline__ float row_pass(const PathParams &prm, uint32_t p_lo, uint32_t p_hi, int my_step,
                                          bool active, float &carry_l, float (&a)[SPL])
{
    return row_finish<SPL, LPR>(prm, row_words<SPL>(prm, p_lo, p_hi, my_step), active, carry_l, a);
}

// ------------------------------------------------------------------------------------------
// General kernel: any n_steps (several passes per row), optional barrier counts and log2 prices.
// SPL = steps per lane (a multiple of 4: one Philox block per 4 steps), LPR = lanes per row
// (32 or 16: a warp walks 32/LPR rows side by side).  logs (nullable): log2 of every stored
// price, the exact FP32 state nested_kernel restarts its inner paths from.
// ------------------------------------------------------------------------------------------
template <int SPL, int LPR, int STORE, bool COUNTS>
__global__ void __launch_bounds__(kPathWarps * 32)
trajectory_kernel(const __grid_constant__ PathParams prm, float *__restrict__ prices, int *__restrict__ counts,
                  float *__restrict__ logs)
{
    constexpr int kBlocks = SPL / 4;
    constexpr int kRowsPerWarp = 32 / LPR;             // rows walked side by side
    constexpr int kPassSteps = SPL * LPR;              // steps of a row covered per pass
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, ln = lane % LPR;       // which of the side-by-side rows, lane within it
    const int n_steps = prm.n_steps;
    // launch-local row indices fit 32 bits (the host caps a launch at 2^31 rows)
    const uint32_t n_rows = (uint32_t)prm.n_paths;
    const uint32_t row_first = (blockIdx.x * kPathWarps + warp) * (kPathsPerWarp * kRowsPerWarp);
    const int lane_step = SPL * ln;

#pragma unroll 1
    for (int r = 0; r < kPathsPerWarp; ++r) {
        if (row_first + (uint32_t)(r * kRowsPerWarp) >= n_rows) break;   // warp-uniform
        const uint32_t row = row_first + (uint32_t)(r * kRowsPerWarp + sub);
        const bool row_ok = row < n_rows;               // a ragged last group computes and discards
        const uint64_t p = prm.first_path + row;
        const uint32_t p_lo = (uint32_t)p, p_hi = (uint32_t)(p >> 32);
        const uint64_t row_off = (uint64_t)row * (uint32_t)n_steps + (uint32_t)lane_step;
        float *out_p = prices + row_off;
        int *out_c = COUNTS ? counts + row_off : nullptr;
        float *out_l = logs ? logs + row_off : nullptr;
        float carry_l = prm.l0;
        int carry_c = 0;

#pragma unroll 1
        for (int step0 = 0; step0 < n_steps; step0 += kPassSteps) {
            const int my_step = step0 + lane_step;
            const bool active = my_step < n_steps;
            float a[SPL];
            const float base = row_pass<SPL, LPR>(prm, p_lo, p_hi, my_step, active, carry_l, a);

            int cbase = 0;
            if (COUNTS) {
                int run = 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) run += (base + a[j] < prm.lB && my_step + j < n_steps) ? 1 : 0;
                cbase = carry_c + group_exclusive_scan<LPR>(run);
                carry_c = __shfl_sync(kFullMask, cbase + run, LPR - 1, LPR);
            }

            if (active && row_ok) {
#pragma unroll
                for (int b = 0; b < kBlocks; ++b) {
                    float l[4], s[4];
                    int c[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        l[j] = base + a[4 * b + j];
                        s[j] = mufu_ex2(l[j]);
                        if (COUNTS) {
                            cbase += (l[j] < prm.lB) ? 1 : 0;
                            c[j] = cbase;
                        }
                    }
                    if (STORE == kStoreVec4) {  // n_steps % 4 == 0: whole float4s are in range
                        if (b == 0 || my_step + 4 * b < n_steps) {
                            __stcs(reinterpret_cast<float4 *>(out_p + step0) + b, make_float4(s[0], s[1], s[2], s[3]));
                            if (COUNTS)
                                __stcs(reinterpret_cast<int4 *>(out_c + step0) + b, make_int4(c[0], c[1], c[2], c[3]));
                            if (logs)
                                __stcs(reinterpret_cast<float4 *>(out_l + step0) + b, make_float4(l[0], l[1], l[2], l[3]));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (my_step + 4 * b + j < n_steps) {
                                __stcs(out_p + step0 + 4 * b + j, s[j]);
                                if (COUNTS) __stcs(out_c + step0 + 4 * b + j, c[j]);
                                if (logs) __stcs(out_l + step0 + 4 * b + j, l[j]);
                            }
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Slab kernel (rows of at most SPL*LPR steps, 16-byte aligned rows; optional barrier counts and
// log2 prices): the bandwidth path of BASELINE config 3 (2^20 x 252).
// A lane's SPL*4 bytes are whole 32-byte sectors only when the row starts on a sector, and with
// 1008-byte rows every other row does not; measured (tools/store_probe.cu) a warp storing 32 B
// per lane reaches 2.9-3.9 TB/s, fully coalesced lines 7.3 TB/s.  So the lanes park their floats
// in shared memory (STS.128), and because a warp walks ROWS CONSECUTIVE rows, which are one
// contiguous slab of ROWS*n_steps*4 bytes in the path-major output, one elected lane hands the
// whole slab to the TMA engine as a single cp.async.bulk (shared::cta -> global): full lines on
// the L1->L2 crossbar, no per-row address arithmetic, one fence per slab.
// Dynamic shared memory: WARPS * arrays * ROWS * n_steps floats.
// ------------------------------------------------------------------------------------------
// ALIGNED: n_steps % 4 == 0, every row is 16-byte aligned and the lanes stage with STS.128.
// Otherwise (150- or 250-step rows ...) the lanes stage element by element, ROWS is a multiple
// of 4 so that every SLAB still starts on a 16-byte boundary of the output, the bulk store takes
// the slab's whole 16-byte units and one lane writes the last one to three floats of a ragged slab.
// FAST (one-pass, 16-byte aligned rows): the lane-invariant Philox products hoisted out of the row loop
// (RowHoist: 16.5 instead of 18.75 IMAD.WIDE per block) and Box-Muller + the base add in packed FP32x2
// instructions (28 fewer issue slots per pass).  Same bits as the plain form.  Measured on 2^20 x 252
// (profiles/r2_trajectory_tuning.txt): prices 250 -> 238 us (4.44 TB/s), prices + counts 354 -> 344 us
// (6.14 TB/s).  (Conflict-free staging -- TMA's 64-byte swizzle + one cp.async.bulk.tensor store per
// slab -- was also built and measured: shared-memory bank conflicts 25.9 M -> 0.9 M, wavefronts 36.9 M ->
// 11.8 M, and NO gain: the conflicts were never the limiter; that variant is not shipped.)
template <int SPL, int LPR, int ROWS, int WARPS, bool COUNTS, bool LOGS, bool MULTI = false, bool ALIGNED = true,
          bool FAST = false>
__global__ void __launch_bounds__(WARPS * 32)
trajectory_slab_kernel(const __grid_constant__ PathParams prm, float *__restrict__ prices, int *__restrict__ counts,
                       float *__restrict__ logs)
{
    static_assert(!FAST || (ALIGNED && !MULTI), "the hoisted products assume one pass per row, the packed stores aligned rows");
    constexpr int kBlocks = SPL / 4;
    constexpr int kRowsPerWarp = 32 / LPR;
    constexpr int kArrays = 1 + (COUNTS ? 1 : 0) + (LOGS ? 1 : 0);
    static_assert(ROWS % kRowsPerWarp == 0, "a slab is a whole number of passes");
    static_assert(ALIGNED || ROWS % 4 == 0, "unaligned rows: slabs must start on 16-byte boundaries");
    extern __shared__ __align__(128) float stage[];    // [warp][array][ROWS][n_steps]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, ln = lane % LPR;
    const int n_steps = prm.n_steps;
    const uint32_t n_rows = (uint32_t)prm.n_paths;
    const uint32_t n_slabs = (n_rows + ROWS - 1) / ROWS;
    const uint32_t slab_stride = gridDim.x * WARPS;   // a warp strides over the slabs (one each when the grid covers them)
    const int lane_step = SPL * ln;
    const int slab_floats = ROWS * n_steps;
    float *my_stage = stage + (size_t)warp * kArrays * slab_floats;
    float *dst0 = my_stage + sub * n_steps + lane_step;
    RowHoist<SPL> hoist;
    if (FAST) hoist = make_row_hoist<SPL>(prm, (uint32_t)(prm.first_path >> 32), lane_step);

#pragma unroll 1
    for (uint32_t slab = blockIdx.x * WARPS + warp; slab < n_slabs; slab += slab_stride) {
        const uint32_t slab_row = slab * ROWS;
        // (Drawing the next pass's Philox blocks while this pass's Box-Muller / exp2 work is in
        // flight was tried -- software pipelining -- and lost 15 % to register pressure.)
#pragma unroll 1
        for (int r = 0; r < ROWS; r += kRowsPerWarp) {
            const uint64_t p = prm.first_path + slab_row + (uint32_t)(r + sub);   // rows past n_rows: computed, not copied
            float carry_l = prm.l0;
            int carry_c = 0;
            // MULTI: rows longer than one pass (SPL*LPR steps) take several, the running log2 price
            // and barrier count carried in registers; otherwise exactly one trip
#pragma unroll 1
            for (int step0 = 0; step0 < (MULTI ? n_steps : 1); step0 += SPL * LPR) {
                const int my_step = step0 + lane_step;
                const bool active = my_step < n_steps;
                PassWords<SPL> words;
                if (FAST && (uint32_t)(p >> 32) == hoist.p_hi)
                    words = row_words_hoisted<SPL>(prm, hoist, (uint32_t)p);
                else words = row_words<SPL>(prm, (uint32_t)p, (uint32_t)(p >> 32), my_step);
                float a[SPL];
                const float base = row_finish<SPL, LPR, FAST>(prm, words, active, carry_l, a);
                if (FAST) {                                                  // log2 prices of this lane's steps
                    const uint64_t bb = f2_pack(base, base);
#pragma unroll
                    for (int j = 0; j < SPL; j += 2) f2_unpack(f2_add(f2_pack(a[j], a[j + 1]), bb), a[j], a[j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < SPL; ++j) a[j] = base + a[j];
                }
                int cbase = carry_c;
                if (COUNTS) {
                    int run = 0;
#pragma unroll
                    for (int j = 0; j < SPL; ++j) run += (a[j] < prm.lB && my_step + j < n_steps) ? 1 : 0;
                    cbase += group_exclusive_scan<LPR>(run);
                    if (MULTI) carry_c = __shfl_sync(kFullMask, cbase + run, LPR - 1, LPR);
                }
                if (r == 0 && step0 == 0) {
                    // the previous slab's bulk copies had this whole pass to read the buffers: wait is ~free
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                float *dst = dst0 + r * n_steps + step0;
#pragma unroll
                for (int b = 0; b < kBlocks; ++b) {
                    if (my_step + 4 * b < n_steps) {
                        int c[4];
                        float s4[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            s4[j] = mufu_ex2(a[4 * b + j]);
                            if (COUNTS) {
                                cbase += (a[4 * b + j] < prm.lB) ? 1 : 0;
                                c[j] = cbase;
                            }
                        }
                        if (ALIGNED) {
                            if (COUNTS)
                                *reinterpret_cast<int4 *>(dst + slab_floats + 4 * b) = make_int4(c[0], c[1], c[2], c[3]);
                            if (LOGS)
                                *reinterpret_cast<float4 *>(dst + (COUNTS ? 2 : 1) * slab_floats + 4 * b) =
                                    make_float4(a[4 * b], a[4 * b + 1], a[4 * b + 2], a[4 * b + 3]);
                            *reinterpret_cast<float4 *>(dst + 4 * b) = make_float4(s4[0], s4[1], s4[2], s4[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (my_step + 4 * b + j < n_steps) {
                                    dst[4 * b + j] = s4[j];
                                    if (COUNTS) reinterpret_cast<int *>(dst + slab_floats)[4 * b + j] = c[j];
                                    if (LOGS) dst[(COUNTS ? 2 : 1) * slab_floats + 4 * b + j] = a[4 * b + j];
                                }
                            }
                        }
                    }
                }
            }
        }
        fence_async_smem();   // generic-proxy STS -> visible to the async proxy (TMA)
        __syncwarp();
        if (lane == 0) {
            const uint32_t rows = min((uint32_t)ROWS, n_rows - slab_row);
            const uint32_t floats = rows * (uint32_t)n_steps;
            const uint32_t bytes = ALIGNED ? floats * 4u : (floats * 4u) & ~15u;   // whole 16-byte units
            const uint64_t off = (uint64_t)slab_row * (uint32_t)n_steps;
            if (bytes) {
                bulk_store(prices + off, my_stage, bytes);
                if (COUNTS) bulk_store(counts + off, my_stage + slab_floats, bytes);
                if (LOGS) bulk_store(logs + off, my_stage + (COUNTS ? 2 : 1) * slab_floats, bytes);
            }
            bulk_commit();
            if (!ALIGNED) {   // a ragged last slab can end in one to three floats past the last unit
                for (uint32_t i = bytes / 4u; i < floats; ++i) {
                    prices[off + i] = my_stage[i];
                    if (COUNTS) counts[off + i] = reinterpret_cast<const int *>(my_stage + slab_floats)[i];
                    if (LOGS) logs[off + i] = my_stage[(COUNTS ? 2 : 1) * slab_floats + i];
                }
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();   // shared memory must outlive the last copies' reads
}

// ------------------------------------------------------------------------------------------
// Long rows (more steps than a slab can stage whole; 16-byte aligned): the same 16 x 16 passes, staged PASS
// BY PASS.  A pass of a row pair is two contiguous 1 KiB pieces of the output per array; the lanes park them
// in one of two shared-memory buffers and an elected lane hands each piece to the TMA engine
// (cp.async.bulk), so the next pass computes while the previous one drains -- instead of 64-byte
// per-lane STG.128 straight from registers (half-filled sectors whenever a row does not start on a sector:
// the general kernel reaches 3.5 TB/s with prices only and 1.4 TB/s with counts at 2048+ steps).
// Same row_words / row_finish as every other trajectory kernel: same bits.
// Dynamic shared memory: WARPS * 2 buffers * arrays * 2 rows * 256 floats.
// ------------------------------------------------------------------------------------------
template <bool COUNTS, bool LOGS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
trajectory_long_kernel(const __grid_constant__ PathParams prm, float *__restrict__ prices, int *__restrict__ counts,
                       float *__restrict__ logs)
{
    constexpr int SPL = 16, LPR = 16, kPass = SPL * LPR, kArrays = 1 + (COUNTS ? 1 : 0) + (LOGS ? 1 : 0);
    constexpr int kBuf = kArrays * 2 * kPass;             // floats of one buffer: [array][row of the pair][256]
    extern __shared__ __align__(128) float stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, ln = lane % LPR;
    const int n_steps = prm.n_steps;
    const uint32_t n_rows = (uint32_t)prm.n_paths;
    const uint32_t n_pairs = (n_rows + 1) / 2;
    const int lane_step = SPL * ln;
    float *my = stage + (size_t)warp * 2 * kBuf;
    unsigned pass_no = 0;                                   // alternates the two buffers across passes AND row pairs

#pragma unroll 1
    for (uint32_t pair = blockIdx.x * WARPS + warp; pair < n_pairs; pair += gridDim.x * WARPS) {
        const uint32_t row = 2 * pair + (uint32_t)sub;      // a row past n_rows is computed and not copied
        const uint64_t p = prm.first_path + row;
        float carry_l = prm.l0;
        int carry_c = 0;
#pragma unroll 1
        for (int step0 = 0; step0 < n_steps; step0 += kPass, ++pass_no) {
            const int my_step = step0 + lane_step;
            const bool active = my_step < n_steps;
            float a[SPL];
            const float base = row_finish<SPL, LPR>(prm, row_words<SPL>(prm, (uint32_t)p, (uint32_t)(p >> 32), my_step),
                                                    active, carry_l, a);
#pragma unroll
            for (int j = 0; j < SPL; ++j) a[j] = base + a[j];
            int cbase = carry_c;
            if (COUNTS) {
                int run = 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) run += (a[j] < prm.lB && my_step + j < n_steps) ? 1 : 0;
                cbase += group_exclusive_scan<LPR>(run);
                carry_c = __shfl_sync(kFullMask, cbase + run, LPR - 1, LPR);
            }
            float *buf = my + (pass_no & 1u) * kBuf;
            if (lane == 0) bulk_wait_read<1>();             // the copies of two passes ago have read this buffer
            __syncwarp();
            float *dst = buf + sub * kPass + lane_step;
#pragma unroll
            for (int b = 0; b < SPL / 4; ++b) {
                if (my_step + 4 * b < n_steps) {            // n_steps % 4 == 0: whole float4s are in range
                    float s4[4];
                    int c[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s4[j] = mufu_ex2(a[4 * b + j]);
                        if (COUNTS) {
                            cbase += (a[4 * b + j] < prm.lB) ? 1 : 0;
                            c[j] = cbase;
                        }
                    }
                    *reinterpret_cast<float4 *>(dst + 4 * b) = make_float4(s4[0], s4[1], s4[2], s4[3]);
                    if (COUNTS) *reinterpret_cast<int4 *>(dst + 2 * kPass + 4 * b) = make_int4(c[0], c[1], c[2], c[3]);
                    if (LOGS)
                        *reinterpret_cast<float4 *>(dst + (COUNTS ? 4 : 2) * kPass + 4 * b) =
                            make_float4(a[4 * b], a[4 * b + 1], a[4 * b + 2], a[4 * b + 3]);
                }
            }
            fence_async_smem();                             // generic-proxy STS -> visible to the async proxy (TMA)
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)min(kPass, n_steps - step0) * 4u;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const uint32_t r = 2 * pair + (uint32_t)s;
                    if (r < n_rows) {
                        const uint64_t off = (uint64_t)r * (uint32_t)n_steps + (uint32_t)step0;
                        bulk_store(prices + off, buf + s * kPass, bytes);
                        if (COUNTS) bulk_store(counts + off, buf + 2 * kPass + s * kPass, bytes);
                        if (LOGS) bulk_store(logs + off, buf + (COUNTS ? 4 : 2) * kPass + s * kPass, bytes);
                    }
                }
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();                     // shared memory must outlive the last copies' reads
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
    float sc, inv_sc, bq, dr;   // sc, dr: see WalkParams; inv_sc = 1 / sc; bq = dr / sc (+inf without a barrier)
    float lB, K;                // log2 B (-inf without a barrier)
    int P1, P2, n_steps, n_inner;
    int discount_mode;          // MCB_DISCOUNT_*
    float r, T, dt;
    uint64_t first_outer;
    PhiloxKeys keys_inner;
};

__global__ void __launch_bounds__(kSlots)   // 55 registers, 4 CTAs per SM (capping it at 48 for a fifth CTA: 51.5 -> 55.9 ms)
nested_kernel(const __grid_constant__ NestedParams prm, const float *__restrict__ logs,
              const int *__restrict__ counts, float *__restrict__ F)
{
    __shared__ float scratch[2 * kWarps];
    extern __shared__ __align__(16) float walk_thr[];   // thr[k] = -(k + 1) bq: the same for every point (see WalkParams)
    const uint64_t p = prm.first_outer + blockIdx.x;
    const int n_steps = prm.n_steps;
    const uint64_t row = (uint64_t)blockIdx.x * (uint64_t)n_steps;
    const bool barrier = prm.lB > -INFINITY;
    fill_walk_thresholds(walk_thr, n_steps, prm.bq);

    // gridDim.y CTAs share an outer trajectory, taking its points k = y, y + gridDim.y, ...: F[p,k]
    // depends on (p, k) only, and the interleaved split balances the work (a point costs
    // N_STEPS - 1 - k inner steps) while giving the scheduler more, shorter CTAs for the tail
#pragma unroll 1
    for (int k = (int)blockIdx.y; k < n_steps; k += (int)gridDim.y) {
        const float lo = __ldg(logs + row + k);
        const int co = __ldg(counts + row + k);
        const int remaining = n_steps - (k + 1);
        // the inner walks of this point start at acc = -(log2 B - lo) / sc and end at l_end + sc * acc
        const float acc0 = barrier ? (lo - prm.lB) * prm.inv_sc : 0.0f;
        const float l_end = fmaf((float)remaining, prm.dr, barrier ? prm.lB : lo);
        float sum = 0.0f, sq = 0.0f;
        if (co <= prm.P2) {
            const uint64_t q = (p * (uint64_t)n_steps + (uint64_t)k) * (uint64_t)prm.n_inner;
            // the point's inner streams q .. q + n_inner - 1 share their high word unless they straddle a
            // multiple of 2^32 (CTA-uniform test): then the round-1 product that depends on it is computed
            // once per warp on the uniform datapath instead of per thread on the multiplier pipe
            const bool hi_uniform = (uint32_t)q + (uint32_t)(prm.n_inner - 1) >= (uint32_t)q;
            const uint32_t q_hi = (uint32_t)(q >> 32);
            int jj = threadIdx.x;
#pragma unroll 1
            for (; jj + kSlots < prm.n_inner; jj += 2 * kSlots) {   // two inner paths interleaved
                const uint64_t sa = q + (uint64_t)jj, sb = sa + kSlots;
                float acc[2] = {acc0, acc0};
                int c[2] = {co, co};
                const uint32_t s_lo[2] = {(uint32_t)sa, (uint32_t)sb};
                if (hi_uniform) {
                    const uint32_t s_hi[2] = {q_hi, q_hi};
                    walk_paths<2>(acc, c, s_lo, s_hi, remaining, walk_thr, prm.bq, prm.keys_inner);
                } else {
                    const uint32_t s_hi[2] = {(uint32_t)(sa >> 32), (uint32_t)(sb >> 32)};
                    walk_paths<2>(acc, c, s_lo, s_hi, remaining, walk_thr, prm.bq, prm.keys_inner);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float pay = (c[i] >= prm.P1 && c[i] <= prm.P2)
                                          ? fmaxf(mufu_ex2(fmaf(prm.sc, acc[i], l_end)) - prm.K, 0.0f) : 0.0f;
                    sum = sum + pay;
                    sq = fmaf(pay, pay, sq);
                }
            }
            if (jj < prm.n_inner) {
                const uint64_t sub = q + (uint64_t)jj;
                float acc = acc0;
                int c = co;
                walk_path(acc, c, (uint32_t)sub, (uint32_t)(sub >> 32), remaining, walk_thr, prm.bq, prm.keys_inner);
                const float pay = (c >= prm.P1 && c <= prm.P2) ? fmaxf(mufu_ex2(fmaf(prm.sc, acc, l_end)) - prm.K, 0.0f)
                                                                : 0.0f;
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
