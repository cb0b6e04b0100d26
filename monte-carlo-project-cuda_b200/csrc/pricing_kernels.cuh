// pricing_kernels.cuh -- the GBM hot path: European, bullet, reductions.  sm_100a only.
//
// Replaces (reference file:line, relative to the reference repo)
//   simulateOptionPriceMultipleBlockGPUwithReduce      inc/trajectories.cuh:54-113
//   simulateBulletOptionPriceMultipleBlockGPU[atomic]  inc/trajectories.cuh:115-271
//   reduce3..6 + host/atomic final sums                inc/reduce.cuh, inc/wrappers.cuh:81-84
// Design: no RNG state, many paths per thread, payoff (sum, sumsq) accumulated in
// registers, one chunk partial per CTA, fixed-order double-precision final passes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "block_reduce.cuh"
#include "philox.cuh"

namespace mcb {

enum { kCall = 0, kPut = 1 };

// ------------------------------------------------------------------------------------------
// European option, one GBM step, canonical (seed, path id) keying.
//   St = S0 exp((r - sigma^2/2) T + sigma sqrt(T) G)  ->  St = 2^(c0 + c1 z)
// with c0 = log2 S0 + (r - sigma^2/2) T log2 e and c1 = sigma sqrt(T) log2 e sqrt(2 ln 2) (the
// Box-Muller radius comes unscaled, philox.cuh) folded on the host, so a path is:
// Philox (block 0) -> Box-Muller sin branch -> FFMA -> MUFU.EX2 -> payoff.
// ------------------------------------------------------------------------------------------
struct EuropeanParams {
    float c0, c1, K;
    uint32_t pad;
    uint64_t n_paths;      // total paths of the run (ragged tail lives in the last chunk)
    uint64_t first_chunk;  // chunk index of blockIdx.x == 0
    PhiloxKeys keys;
};

// Philox block 0 of path p with the first round peeled: the counter is (0, 0, p_lo, p_hi), so
// round 1 needs only prod1 = M1 * p_lo -- and consecutive paths of a slot differ by 256, so the
// caller advances prod1 with a 64-bit add (ALU pipe) instead of re-multiplying (fmaheavy pipe,
// the kernel's bottleneck).  Same words as philox4x32_10(0, 0, p_lo, p_hi).
__device__ __forceinline__ Words4 philox_block0_from_prod(uint64_t prod1, uint32_t p_hi, const PhiloxKeys &k)
{
    uint32_t c0 = (uint32_t)(prod1 >> 32) ^ k.k0[0];
    uint32_t c1 = (uint32_t)prod1;
    uint32_t c2 = p_hi ^ k.k1[0];
    uint32_t c3 = 0u;
#pragma unroll
    for (int r = 1; r < 10; ++r) philox_round(c0, c1, c2, c3, k.k0[r], k.k1[r]);
    return Words4{c0, c1, c2, c3};
}

template <int TYPE>
__device__ __forceinline__ float european_payoff_from_prod(uint64_t prod1, uint32_t p_hi, const EuropeanParams &prm)
{
    const Words4 w = philox_block0_from_prod(prod1, p_hi, prm.keys);
    const float St = mufu_ex2(fmaf(prm.c1, unit_normal_sin(w.x, w.y), prm.c0));
    return TYPE == kPut ? fmaxf(prm.K - St, 0.0f) : fmaxf(St - prm.K, 0.0f);
}

template <int TYPE>
__device__ __forceinline__ float european_payoff(uint32_t p_lo, uint32_t p_hi, const EuropeanParams &prm)
{
    return european_payoff_from_prod<TYPE>((uint64_t)kPhiloxM1 * p_lo, p_hi, prm);
}

// One CTA = one chunk of kSlots * PPS consecutive paths.  Slot t (= thread t) accumulates
// chunk-local paths t, t+256, ... in that order; partials[chunk - first_chunk] = (sum, sumsq).
// payoffs (nullable) receives every path's payoff (parity hook).
template <int TYPE, int PPS, int UNROLL = 4>
__global__ void __launch_bounds__(kSlots)
european_kernel(const __grid_constant__ EuropeanParams prm, float2 *__restrict__ partials,
                float *__restrict__ payoffs, uint64_t payoffs_first_path)
{
    __shared__ float scratch[2 * kWarps];
    const uint64_t chunk = prm.first_chunk + blockIdx.x;
    const uint64_t base = chunk * (uint64_t)(kSlots * PPS);
    // a chunk never straddles a multiple of 2^32, so the high counter word is CTA-uniform
    const uint32_t p_hi = (uint32_t)(base >> 32);
    const uint32_t p_lo0 = (uint32_t)base + threadIdx.x;
    const uint64_t left = prm.n_paths - base;  // > 0 by construction of the grid

    float sum = 0.0f, sq = 0.0f;
    if (left >= (uint64_t)(kSlots * PPS) && payoffs == nullptr) {
        uint64_t prod1 = (uint64_t)kPhiloxM1 * p_lo0;   // a chunk never wraps p_lo, so this stays M1 * p_lo
#pragma unroll UNROLL
        for (int i = 0; i < PPS; ++i) {
            const float pay = european_payoff_from_prod<TYPE>(prod1, p_hi, prm);
            prod1 += (uint64_t)kPhiloxM1 * kSlots;
            sum = sum + pay;
            sq = fmaf(pay, pay, sq);
        }
    } else {
        for (int i = 0; i < PPS; ++i) {
            const uint32_t local = (uint32_t)(i * kSlots) + threadIdx.x;
            if ((uint64_t)local < left) {
                const float pay = european_payoff<TYPE>(p_lo0 + (uint32_t)(i * kSlots), p_hi, prm);
                sum = sum + pay;
                sq = fmaf(pay, pay, sq);
                if (payoffs && base + local >= payoffs_first_path) payoffs[base + local - payoffs_first_path] = pay;
            }
        }
    }
    block_fold2(sum, sq, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(sum, sq);
}

// ------------------------------------------------------------------------------------------
// European option under PACKED keying (SURVEY.md 8(d) "packed keying", reported separately from the canonical
// figure): path p draws normal p & 3 of the stream (seed, subsequence p >> 2) -- the four values four successive
// curand_normal() calls return -- so ONE Philox block prices FOUR paths: 4.25 IMAD.WIDE and 3 MUFU per path
// instead of 16 and 4, and the bound moves from the multiplier pipe to the XU pipe (16/3 paths/clk/SM).
// One CTA = one chunk of kSlots * PPS paths = kSlots * PPS / 4 blocks; slot t takes the chunk's blocks
// t, t + 256, ... and accumulates each block's four paths in order, then the usual block tree.
// ------------------------------------------------------------------------------------------
template <int TYPE>
__device__ __forceinline__ void packed_payoffs4(const Words4 &w, const EuropeanParams &prm, float pay[4])
{
    const float t0 = prm.c1 * bm_radius_unscaled(w.x), t1 = prm.c1 * bm_radius_unscaled(w.z);
    const float v0 = bm_angle(w.y), v1 = bm_angle(w.w);
    const float l[4] = {fmaf(t0, mufu_sin(v0), prm.c0), fmaf(t0, mufu_cos(v0), prm.c0),
                        fmaf(t1, mufu_sin(v1), prm.c0), fmaf(t1, mufu_cos(v1), prm.c0)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float St = mufu_ex2(l[j]);
        pay[j] = TYPE == kPut ? fmaxf(prm.K - St, 0.0f) : fmaxf(St - prm.K, 0.0f);
    }
}

template <int TYPE, int PPS>
__global__ void __launch_bounds__(kSlots)
european_packed_kernel(const __grid_constant__ EuropeanParams prm, float2 *__restrict__ partials,
                       float *__restrict__ payoffs, uint64_t payoffs_first_path)
{
    static_assert(PPS % 4 == 0, "a slot takes whole Philox blocks");
    __shared__ float scratch[2 * kWarps];
    const uint64_t chunk = prm.first_chunk + blockIdx.x;
    const uint64_t base = chunk * (uint64_t)(kSlots * PPS);          // first path of the chunk
    const uint64_t sbase = base >> 2;                                 // its first subsequence (= block of four paths)
    // a chunk's kSlots * PPS / 4 subsequences never straddle a multiple of 2^32: the high word is CTA-uniform
    const uint32_t s_hi = (uint32_t)(sbase >> 32);
    const uint32_t s_lo0 = (uint32_t)sbase + threadIdx.x;
    const uint64_t left = prm.n_paths - base;                         // > 0 by construction of the grid

    float sum = 0.0f, sq = 0.0f;
    if (left >= (uint64_t)(kSlots * PPS) && payoffs == nullptr) {
        uint64_t prod1 = (uint64_t)kPhiloxM1 * s_lo0;                 // round 0's only product, advanced by an add
#pragma unroll 2
        for (int i = 0; i < PPS / 4; ++i) {
            float pay[4];
            packed_payoffs4<TYPE>(philox_block0_from_prod(prod1, s_hi, prm.keys), prm, pay);
            prod1 += (uint64_t)kPhiloxM1 * kSlots;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sum = sum + pay[j];
                sq = fmaf(pay[j], pay[j], sq);
            }
        }
    } else {
        for (int i = 0; i < PPS / 4; ++i) {
            const uint32_t blk = (uint32_t)(i * kSlots) + threadIdx.x;     // chunk-local block
            if ((uint64_t)blk * 4 < left) {
                float pay[4];
                packed_payoffs4<TYPE>(philox4x32_10(0u, 0u, s_lo0 + (uint32_t)(i * kSlots), s_hi, prm.keys), prm, pay);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t local = (uint64_t)blk * 4 + (uint64_t)j;
                    if (local < left) {
                        sum = sum + pay[j];
                        sq = fmaf(pay[j], pay[j], sq);
                        if (payoffs && base + local >= payoffs_first_path) payoffs[base + local - payoffs_first_path] = pay[j];
                    }
                }
            }
        }
    }
    block_fold2(sum, sq, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(sum, sq);
}

// ------------------------------------------------------------------------------------------
// Batched strike/vol sweep with common random numbers (BASELINE config 5): every parameter set
// is priced on the SAME draws, so the Philox + Box-Muller work is done once per path and only
//   FFMA, MUFU.EX2, FADD, FMNMX, FADD, FFMA
// is repeated per (path, parameter set) -- the kernel is MUFU-bound (one EX2 per unit).
// One CTA = one chunk; a thread keeps the 64 unit normals of its slot in registers and walks
// the parameter sets, folding each set's (sum, sumsq) through the same slot order and the same
// block tree as european_kernel: partials[set][chunk] is BIT-identical to a separate
// european_kernel launch with that set's (K, sigma).
// ------------------------------------------------------------------------------------------
struct SweepParams {
    uint64_t n_paths;      // total paths of the run
    uint64_t first_chunk;  // chunk index of blockIdx.x == 0
    uint64_t stride;       // float2 elements between consecutive parameter sets in `partials`
    int n_sets;
    int sets_per_cta;      // blockIdx.y walks the sets [y*sets_per_cta, (y+1)*sets_per_cta): more, shorter CTAs
    PhiloxKeys keys;
};

constexpr int kSweepTile = 4;  // parameter sets folded together (8 values through one tree)

template <int TYPE, int PPS>
__global__ void __launch_bounds__(kSlots)
sweep_kernel(const __grid_constant__ SweepParams prm, const float4 *__restrict__ sets /* (c0, c1, K, -) */,
             float2 *__restrict__ partials)
{
    __shared__ float scratch[2][2 * kSweepTile][kWarps];   // double-buffered: one barrier per tile
    int tile_parity = 0;
    static_assert(kSweepTile == 4 && kWarps == 8, "the tile fold below is written for 8 values x 8 warps");
    const uint64_t chunk = prm.first_chunk + blockIdx.x;
    const uint64_t base = chunk * (uint64_t)(kSlots * PPS);
    const uint32_t p_hi = (uint32_t)(base >> 32);
    const uint32_t p_lo0 = (uint32_t)base + threadIdx.x;
    const uint64_t left = prm.n_paths - base;
    // paths of this slot that exist: chunk-local indices t, t+256, ... < left
    const int n_valid = left >= (uint64_t)(kSlots * PPS)
                            ? PPS
                            : (int)((left + (uint64_t)(kSlots - 1) - threadIdx.x) / kSlots);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float unit[PPS];  // normal / sqrt(2 ln 2) of every path of the slot
#pragma unroll
    for (int i = 0; i < PPS; ++i) {
        const Words4 w = philox4x32_10(0u, 0u, p_lo0 + (uint32_t)(i * kSlots), p_hi, prm.keys);
        unit[i] = unit_normal_sin(w.x, w.y);
    }

    // splitting the sets over blockIdx.y repeats the draw (57 instructions per path against 7 per
    // (path, set)) but keeps the grid at many waves when a rank owns few chunks
    const int set_begin = (int)blockIdx.y * prm.sets_per_cta;
    const int set_end = min(prm.n_sets, set_begin + prm.sets_per_cta);
    for (int s0 = set_begin; s0 < set_end; s0 += kSweepTile) {
        float sum[kSweepTile], sq[kSweepTile];
        // Two parameter sets ride in one packed FP32x2 register pair: per (path, pair of sets)
        //   FFMA2, 2 x MUFU.EX2, FADD2, 2 x FMNMX, FADD2, FFMA2
        // = 4 issue slots per (path, set) instead of 6, so the schedulers keep the XU pipe (one EX2 per
        // unit, 8 clk per warp instruction) fed.  Every half rounds like the scalar instruction of
        // european_kernel: same bits.
#pragma unroll
        for (int k = 0; k < kSweepTile; k += 2) {
            const bool have0 = s0 + k < set_end, have1 = s0 + k + 1 < set_end;
            const float4 ca = have0 ? __ldg(sets + s0 + k) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float4 cb = have1 ? __ldg(sets + s0 + k + 1) : ca;
            const uint64_t cx = f2_pack(ca.x, cb.x), cy = f2_pack(ca.y, cb.y), ck = f2_pack(ca.z, cb.z);
            uint64_t acc = f2_pack(0.0f, 0.0f), acq = f2_pack(0.0f, 0.0f);
            auto one_path = [&](float z) {
                float l0, l1, d0, d1;
                f2_unpack(f2_fma(cy, f2_pack(z, z), cx), l0, l1);
                const uint64_t st = f2_pack(mufu_ex2(l0), mufu_ex2(l1));
                f2_unpack(TYPE == kPut ? f2_sub(ck, st) : f2_sub(st, ck), d0, d1);
                const uint64_t pay = f2_pack(fmaxf(d0, 0.0f), fmaxf(d1, 0.0f));
                acc = f2_add(acc, pay);
                acq = f2_fma(pay, pay, acq);
            };
            if (n_valid == PPS) {            // every chunk but a ragged last one: straight-line code
#pragma unroll
                for (int i = 0; i < PPS; ++i) one_path(unit[i]);
            } else {
#pragma unroll
                for (int i = 0; i < PPS; ++i)
                    if (i < n_valid) one_path(unit[i]);
            }
            f2_unpack(acc, sum[k], sum[k + 1]);
            f2_unpack(acq, sq[k], sq[k + 1]);
            if (!have0) sum[k] = sq[k] = 0.0f;
            if (!have1) sum[k + 1] = sq[k + 1] = 0.0f;
        }
        // block_fold2's tree for the tile's eight values (sum, sumsq of four sets) at once.  A plain
        // warp_fold per value costs 8 x 5 shuffles, and shuffles queue in the same MIO pipe as the MUFU
        // instructions this kernel is bound by (ncu: mio_throttle is its top stall).  Here the lanes split the
        // eight values between them as they fold -- at offset 16 each lane keeps four values and receives its
        // partner's four, at offset 8 two, at offset 4 one -- so the whole warp stage is 4 + 2 + 1 + 1 + 1 = 9
        // shuffles.  Every addition pairs exactly the operands warp_fold's lane 0 pairs (slot p with slot
        // p + offset; IEEE addition commutes), so the bits are unchanged; lane 4k ends up with value k.
        const float v8[2 * kSweepTile] = {sum[0], sq[0], sum[1], sq[1], sum[2], sq[2], sum[3], sq[3]};
        float k4[4], k2[2], k1;
        {
            const bool up = (lane & 16) != 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float keep = up ? v8[4 + j] : v8[j], send = up ? v8[j] : v8[4 + j];
                k4[j] = keep + __shfl_xor_sync(kFullMask, send, 16);
            }
        }
        {
            const bool up = (lane & 8) != 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float keep = up ? k4[2 + j] : k4[j], send = up ? k4[j] : k4[2 + j];
                k2[j] = keep + __shfl_xor_sync(kFullMask, send, 8);
            }
        }
        {
            const bool up = (lane & 4) != 0;
            const float keep = up ? k2[1] : k2[0], send = up ? k2[0] : k2[1];
            k1 = keep + __shfl_xor_sync(kFullMask, send, 4);
        }
        k1 = k1 + __shfl_xor_sync(kFullMask, k1, 2);
        k1 = k1 + __shfl_xor_sync(kFullMask, k1, 1);
        float(*buf)[kWarps] = scratch[tile_parity];
        if ((lane & 3) == 0) buf[lane >> 2][warp] = k1;          // value index = lane / 4
        __syncthreads();
        // the 8 -> 1 step over the warps, four values at a time on eight lanes each; the other warps go on to
        // the next tile (its scratch is the other buffer, so one barrier per tile is enough)
        if (warp == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float x = buf[4 * h + (lane >> 3)][lane & 7];
#pragma unroll
                for (int off = kWarps / 2; off > 0; off >>= 1) x = x + __shfl_down_sync(kFullMask, x, off, 8);
                const float q = __shfl_down_sync(kFullMask, x, 8);           // the sum of squares sits 8 lanes up
                const int set = s0 + 2 * h + (lane >> 4);
                if ((lane & 15) == 0 && set < set_end)
                    partials[(uint64_t)set * prm.stride + blockIdx.x] = make_float2(x, q);
            }
        }
        tile_parity ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// Multi-step GBM walk in log2 space with the drift and the volatility taken OUT of the loop.
// After k + 1 steps   log2 S = l0 + (k + 1) dr + sc * sum_{i <= k} z_i   (z_i: unit normals / sqrt(2 ln 2),
// sc = sigma sqrt(dt) log2 e sqrt(2 ln 2), dr = (r - sigma^2/2) dt log2 e), so with
//     acc_k = sum_{i <= k} z_i - A,   A = (log2 B - l0) / sc,   Bq = dr / sc
// the barrier test  log2 S < log2 B  is  acc_k < -(k + 1) Bq:  the right-hand sides are the same for
// every path (a table of n_steps floats in shared memory, one LDS.128 per Philox block), and a step is
//     acc = fma(radius, trig, acc);  count += acc < thr[k]
// -- ONE FFMA per step, where the scaled-increment form needs a multiply, an FFMA and an add (every FP32
// instruction costs this loop ~1 clk of the pipe the Philox multiplies saturate, profiles/r2_pipe_microbench.txt).
// At the end  log2 S_T = l_end + sc * acc  with l_end = log2 B + n dr  (l0 + n dr and acc_0 = 0, thr = -inf
// when there is no barrier).  One Philox block feeds four steps; one MUFU.EX2 per path.
// ------------------------------------------------------------------------------------------
struct WalkParams {
    float acc0;    // -(log2 B - l0) / sc   (0 without a barrier)
    float sc;      // sigma sqrt(dt) log2(e) sqrt(2 ln 2): scale of the unscaled Box-Muller radius (> 0)
    float bq;      // dr / sc: the barrier thresholds are -(k + 1) bq   (+inf without a barrier)
    float l_end;   // log2 B + n_steps dr   (l0 + n_steps dr without a barrier)
    float K;
    int P1, P2;
    int n_steps;   // steps to walk
    int count0;    // initial barrier count (Ik)
    int table;     // 1: the launch carries n_steps (rounded up to 4) floats of dynamic shared memory for thr[]
    uint64_t n_paths;
    uint64_t first_chunk;
    PhiloxKeys keys;
};

// thr[k] = -(k + 1) bq for the steps [0, n_steps), padded to a multiple of 4; all threads of the CTA call it
__device__ __forceinline__ void fill_walk_thresholds(float *thr, int n_steps, float bq)
{
    const int n4 = (n_steps + 3) & ~3;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) thr[i] = -(float)(i + 1) * bq;
    __syncthreads();
}

// Walk `n_steps` steps of N independent streams (keys, subsequence[k]) from (acc[k], count[k]);
// normal i of a stream drives step i.  thr: the threshold table (nullptr: the four thresholds of a block
// are computed on the spot, for walks too long for shared memory).  N > 1 interleaves the integer
// (Philox) chains of several paths in one thread: same arithmetic per path, more instruction-level
// parallelism.
template <int N>
__device__ __forceinline__ void walk_paths(float (&acc)[N], int (&count)[N], const uint32_t (&s_lo)[N],
                                           const uint32_t (&s_hi)[N], int n_steps, const float *thr, float bq,
                                           const PhiloxKeys &keys)
{
    auto one_block = [&](int b, int steps /* 4, or 1..3 in a ragged last block */) {
        float4 th;
        if (thr) {
            th = *reinterpret_cast<const float4 *>(thr + 4 * b);
        } else {
            th.x = -(float)(4 * b + 1) * bq;
            th.y = -(float)(4 * b + 2) * bq;
            th.z = -(float)(4 * b + 3) * bq;
            th.w = -(float)(4 * b + 4) * bq;
        }
        float r[N][2], g[N][4];
#pragma unroll
        for (int k = 0; k < N; ++k) unit_factors4(philox4x32_10((uint32_t)b, 0u, s_lo[k], s_hi[k], keys), r[k], g[k]);
        const float t[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < steps) {
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    acc[k] = fmaf(r[k][j >> 1], g[k][j], acc[k]);
                    count[k] += (acc[k] < t[j]) ? 1 : 0;
                }
            }
        }
    };
    const int full = n_steps >> 2;
#pragma unroll 1
    for (int b = 0; b < full; ++b) one_block(b, 4);
    if (n_steps & 3) one_block(full, n_steps & 3);
}

__device__ __forceinline__ void walk_path(float &acc, int &count, uint32_t s_lo, uint32_t s_hi, int n_steps,
                                          const float *thr, float bq, const PhiloxKeys &keys)
{
    float a1[1] = {acc};
    int c1[1] = {count};
    const uint32_t lo1[1] = {s_lo}, hi1[1] = {s_hi};
    walk_paths<1>(a1, c1, lo1, hi1, n_steps, thr, bq, keys);
    acc = a1[0];
    count = c1[0];
}

__device__ __forceinline__ float bullet_payoff_from(float acc, int count, const WalkParams &prm)
{
    const float pay = fmaxf(mufu_ex2(fmaf(prm.sc, acc, prm.l_end)) - prm.K, 0.0f);
    return (count >= prm.P1 && count <= prm.P2) ? pay : 0.0f;
}

template <int PPS>
__global__ void __launch_bounds__(kSlots, 6)   // 40 registers: 48 warps per SM (the loop needs 42 without the cap)
bullet_kernel(const __grid_constant__ WalkParams prm, float2 *__restrict__ partials,
              float *__restrict__ payoffs, uint64_t payoffs_first_path)
{
    __shared__ float scratch[2 * kWarps];
    extern __shared__ __align__(16) float walk_thr[];
    const float *thr = prm.table ? walk_thr : nullptr;
    if (prm.table) fill_walk_thresholds(walk_thr, prm.n_steps, prm.bq);
    const uint64_t chunk = prm.first_chunk + blockIdx.x;
    const uint64_t base = chunk * (uint64_t)(kSlots * PPS);
    const uint64_t left = prm.n_paths - base;
    float sum = 0.0f, sq = 0.0f;
    static_assert(PPS % 2 == 0, "paths of a slot are walked two at a time");
#pragma unroll 1
    for (int i = 0; i < PPS; i += 2) {
        const uint32_t local = (uint32_t)(i * kSlots) + threadIdx.x;
        if ((uint64_t)local + kSlots < left) {
            // both paths of the pair exist: walk them interleaved, accumulate in slot order
            const uint64_t pa = base + local, pb = pa + kSlots;
            float acc[2] = {prm.acc0, prm.acc0};
            int count[2] = {prm.count0, prm.count0};
            // a chunk (1024 consecutive, 1024-aligned paths) never straddles a multiple of 2^32: the high
            // stream word is the CTA-uniform high word of `base`, so the round-1 product it feeds
            // (M1 * (hi(M0 * block) ^ p_hi ^ k1[0])) lives on the uniform datapath, off the multiplier pipe
            const uint32_t p_hi = (uint32_t)(base >> 32);
            const uint32_t lo[2] = {(uint32_t)pa, (uint32_t)pb}, hi[2] = {p_hi, p_hi};
            walk_paths<2>(acc, count, lo, hi, prm.n_steps, thr, prm.bq, prm.keys);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float pay = bullet_payoff_from(acc[k], count[k], prm);
                sum = sum + pay;
                sq = fmaf(pay, pay, sq);
                const uint64_t p = k ? pb : pa;
                if (payoffs && p >= payoffs_first_path) payoffs[p - payoffs_first_path] = pay;
            }
        } else if ((uint64_t)local < left) {
            const uint64_t p = base + local;
            float acc = prm.acc0;
            int count = prm.count0;
            walk_path(acc, count, (uint32_t)p, (uint32_t)(p >> 32), prm.n_steps, thr, prm.bq, prm.keys);
            const float pay = bullet_payoff_from(acc, count, prm);
            sum = sum + pay;
            sq = fmaf(pay, pay, sq);
            if (payoffs && p >= payoffs_first_path) payoffs[p - payoffs_first_path] = pay;
        }
    }
    block_fold2(sum, sq, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(sum, sq);
}

// ------------------------------------------------------------------------------------------
// Final passes (replace the float atomicAdd / host loop of the reference).
// segment_kernel: CTA s folds the chunk partials of segment s in double, fixed order.
//   partials is indexed by (chunk - partials_first_chunk); segments outside [seg_lo, seg_hi)
//   are written as +0.0 so a sum-allreduce over ranks reproduces every segment exactly
//   (or left alone, write_unowned == 0, when the shards share one destination buffer).
// combine_kernel: one warp folds the 64 segments and finalises price + standard error.
// Both are batched over parameter sets with blockIdx.y / blockIdx.x.
// ------------------------------------------------------------------------------------------
constexpr int kSegments = 64;  // MCB_SEGMENTS

__global__ void __launch_bounds__(kSlots)
segment_kernel(const float2 *__restrict__ partials, uint64_t partials_stride, uint64_t partials_first_chunk,
               uint64_t n_chunks, int seg_lo, int seg_hi, int write_unowned, double *__restrict__ segments)
{
    __shared__ double scratch[2 * kWarps];
    const int seg = blockIdx.x;
    const int set = blockIdx.y;
    double a = 0.0, b = 0.0;
    // write_unowned == 0: `segments` is shared by several shards (the leader's buffer, written over
    // NVLink by the others); every shard writes exactly the segments it owns
    if (!write_unowned && (seg < seg_lo || seg >= seg_hi)) return;   // CTA-uniform
    if (seg >= seg_lo && seg < seg_hi) {
        const uint64_t lo = (n_chunks * (uint64_t)seg) / kSegments;
        const uint64_t hi = (n_chunks * (uint64_t)(seg + 1)) / kSegments;
        const float2 *src = partials + (uint64_t)set * partials_stride;
        for (uint64_t c = lo + threadIdx.x; c < hi; c += kSlots) {
            const float2 v = src[c - partials_first_chunk];
            a = a + (double)v.x;
            b = b + (double)v.y;
        }
    }
    block_fold2(a, b, scratch);
    if (threadIdx.x == 0) {
        double *dst = segments + ((uint64_t)set * kSegments + seg) * 2;
        dst[0] = a;
        dst[1] = b;
    }
}

struct ResultDev {  // == mcb_result
    double price, std_error, sum, sumsq;
    uint64_t n_paths;
};

__global__ void __launch_bounds__(32)
combine_kernel(const double *__restrict__ segments, uint64_t n_paths, double discount, ResultDev *__restrict__ out)
{
    const int set = blockIdx.x, lane = threadIdx.x;
    const double *seg = segments + (uint64_t)set * kSegments * 2;
    double s = seg[2 * lane] + seg[2 * (lane + 32)];
    double q = seg[2 * lane + 1] + seg[2 * (lane + 32) + 1];
    s = warp_fold(s);
    q = warp_fold(q);
    if (lane == 0) {
        const double n = (double)n_paths;
        const double mean = s / n;
        double var = q / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        if (n_paths > 1) var *= n / (n - 1.0);
        ResultDev r;
        r.price = discount * mean;
        r.std_error = discount * sqrt(var / n);
        r.sum = s;
        r.sumsq = q;
        r.n_paths = n_paths;
        out[set] = r;
    }
}

// The same segment pass for MANY parameter sets (the sweep: 1024 sets x 64 segments = 65 536 folds of a few
// dozen partials each): one WARP per (set, segment) instead of one CTA, walking the eight "virtual warps"
// of segment_kernel's 256-slot tree one after the other -- the same additions in the same order, so the
// same bits, with an eighth of the CTAs and no barrier.
__global__ void __launch_bounds__(kSlots)
segment_sets_kernel(const float2 *__restrict__ partials, uint64_t partials_stride, uint64_t partials_first_chunk,
                    uint64_t n_chunks, int seg_lo, int seg_hi, int write_unowned, int n_sets,
                    double *__restrict__ segments)
{
    const int lane = threadIdx.x & 31;
    const uint64_t item = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (item >= (uint64_t)n_sets * kSegments) return;                  // warp-uniform
    const int set = (int)(item / kSegments), seg = (int)(item % kSegments);
    const bool owned = seg >= seg_lo && seg < seg_hi;
    if (!owned && !write_unowned) return;
    double ta = 0.0, tb = 0.0;                                          // lane v < 8: total of virtual warp v
    if (owned) {
        const uint64_t lo = (n_chunks * (uint64_t)seg) / kSegments;
        const uint64_t hi = (n_chunks * (uint64_t)(seg + 1)) / kSegments;
        const float2 *src = partials + (uint64_t)set * partials_stride;
        for (int v = 0; v < kWarps; ++v) {
            if (lo + (uint64_t)(32 * v) >= hi) break;                   // the remaining slots are all +0.0
            double a = 0.0, b = 0.0;
            for (uint64_t c = lo + (uint64_t)(32 * v + lane); c < hi; c += kSlots) {
                const float2 p = src[c - partials_first_chunk];
                a = a + (double)p.x;
                b = b + (double)p.y;
            }
            a = __shfl_sync(kFullMask, warp_fold(a), 0);
            b = __shfl_sync(kFullMask, warp_fold(b), 0);
            if (lane == v) {
                ta = a;
                tb = b;
            }
        }
    }
#pragma unroll
    for (int off = kWarps / 2; off > 0; off >>= 1) {                    // block_fold2's 8 -> 1 step
        ta = ta + __shfl_down_sync(kFullMask, ta, off);
        tb = tb + __shfl_down_sync(kFullMask, tb, off);
    }
    if (lane == 0) {
        double *dst = segments + ((uint64_t)set * kSegments + seg) * 2;
        dst[0] = ta;
        dst[1] = tb;
    }
}

// ------------------------------------------------------------------------------------------
// The European JOB pipeline (mcb_european_submit / mcb_european_collect; one launch per shard).
//
// A job is priced by `world` shards (GPUs), shard g owning the chunks of segments
// [64g/world, 64(g+1)/world).  european_job_kernel is european_kernel's chunk loop plus, through
// tickets and with no second launch:
//   * the last CTA of every segment folds that segment (same order as segment_kernel) and STORES
//     the (sum, sumsq) pair into the mailbox slot of every consumer shard -- its own HBM, or a
//     peer's over NVLink / NVSwitch (plain st.global on a peer-mapped address);
//   * the CTA that completes the shard's last segment writes +0.0 into owned segments that have
//     no chunk, fences and publishes the job's epoch flag in every consumer's mailbox;
//   * with world == 1 that same CTA runs the final tree (same order as combine_kernel) and
//     writes the result to device memory and to a mapped host slot: ONE launch per price.
// With world > 1 the final tree is combine_job_kernel on the consumer's OWN second stream: it
// waits (bounded by a %globaltimer deadline) for the `world` flags, folds, writes the result and
// acknowledges the slot to every producer.  The pricing stream never waits for a peer, so rank
// skew and NVLink latency hide behind the next job's pricing; a ring of kRing mailbox slots
// lets a shard run kRing - 1 jobs ahead of the slowest consumer.
// The cold tail lives in a __noinline__ function so that the hot loop keeps its 8 CTAs per SM.
// ------------------------------------------------------------------------------------------
#ifndef MCB_TRACE                 // tools/job_latency_probe.cu defines it to stamp clock64() along the one-CTA job
#define MCB_TRACE(i)
#endif
#ifndef MCB_TRACE_CTA             // ... and to stamp %globaltimer in every CTA of the small-job kernel
#define MCB_TRACE_CTA(i)
#endif
constexpr int kMaxPeers = 16;  // MCB_MAX_PEERS
constexpr int kRing = 4;       // MCB_PIPELINE_DEPTH: mailbox slots = jobs in flight per engine

struct PeerMailbox {           // one per shard, in that shard's HBM; written by every producer shard
    double gather[kRing][kSegments * 2];               // (sum, sumsq) segments of the job in each slot
    unsigned long long flags[kRing][kMaxPeers];        // flags[slot][r] = last epoch producer r published there
    unsigned long long consumed[kMaxPeers];            // consumed[c] = last epoch consumer c folded (its ack to ME)
    unsigned int timeouts;                             // bounded waits that ran out (sticky, diagnostic)
    // small one-GPU jobs: the segments as self-tagged words (low half | tag << 32, high half | tag << 32 of sum and
    // sumsq), so that no fence and no ticket stand between a chunk's totals and the final tree (small_job_tail_one_gpu)
    alignas(32) unsigned long long tagged[kRing][kSegments][4];   // (read with 16-byte loads)
    double own[kRing][kSegments * 2];                  // ... and in the clear for mcb_last_segments; no peer writes here
};
struct PeerTable {
    PeerMailbox *box[kMaxPeers];                       // box[r] = shard r's mailbox as mapped here
};

// Result slot in mapped pinned host memory, one per job in flight.  The five 8-byte fields of ResultDev travel as ten
// words of (data half | job tag << 32): every 8-byte store is a complete message, so the kernel needs NO system fence
// between "data" and "flag" (fence.sys costs 1.4 us here, tools/launch_floor.cu) and the host accepts the result once
// all ten words carry the job's tag -- the flag-in-data scheme of the low-latency protocols of collective libraries.
struct HostSlot {
    unsigned long long w[10];  // w[2k] = low half of field k, w[2k + 1] = high half, tag = low 32 bits of the epoch
    unsigned long long pad[6];
};

__device__ __forceinline__ void host_publish(HostSlot *h, const ResultDev &r, unsigned long long epoch)
{
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
    const unsigned long long f[5] = {(unsigned long long)__double_as_longlong(r.price),
                                     (unsigned long long)__double_as_longlong(r.std_error),
                                     (unsigned long long)__double_as_longlong(r.sum),
                                     (unsigned long long)__double_as_longlong(r.sumsq), (unsigned long long)r.n_paths};
    volatile unsigned long long *w = h->w;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        w[2 * k] = tag | (f[k] & 0xffffffffull);
        w[2 * k + 1] = tag | (f[k] >> 32);
    }
}

struct JobArgs {
    uint64_t n_chunks;              // chunks of the WHOLE job (segment boundaries are global)
    uint64_t n_paths;
    double discount;
    unsigned int *seg_tickets;      // [kSegments] + [1] (shard ticket), all zero between launches
    PeerTable peers;
    ResultDev *d_out;               // world == 1: device copy of the result (nullable)
    HostSlot *h_out;                // world == 1: mapped host slot (nullable)
    unsigned long long epoch;       // job number, > 0, agreed by every shard
    unsigned long long timeout_ns;  // bound of every device-side wait
    int seg_lo, seg_hi;             // segments this shard owns
    int live_segments;              // ... of which have at least one chunk
    int rank, world;
    int n_consumers;                // shards [0, n_consumers) receive the segments (world: all-gather; 1: gather)
    unsigned long long ack_epoch;   // separate processes: the last SHARDED job that used this mailbox slot; producers wait
                                    // until every consumer has acknowledged it (0: nothing to wait for)
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin (with back-off) until *p >= want or timeout_ns has passed.  The word lives in THIS GPU's
// memory and is written by a peer over NVLink; volatile loads read it at L2.
__device__ __noinline__ bool wait_at_least(const volatile unsigned long long *p, unsigned long long want,
                                           unsigned long long timeout_ns)
{
    if (*p >= want) return true;
    const unsigned long long t0 = global_timer_ns();
    unsigned int ns = 32;
    while (*p < want) {
        if (global_timer_ns() - t0 > timeout_ns) return false;
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
    }
    return true;
}

// The fixed 64 -> 1 tree of combine_kernel, run by one warp on a mailbox slot.  Segments without a
// chunk (jobs of fewer than 64 chunks) are +0.0 by rule and are never stored or read; ok == false (a
// peer never arrived) poisons the result with NaN and n_paths = 0 instead of a plausible number.
__device__ __forceinline__ bool segment_is_empty(uint64_t n_chunks, int seg)
{
    return (n_chunks * (uint64_t)seg) / kSegments == (n_chunks * (uint64_t)(seg + 1)) / kSegments;
}

// (s, q) = this lane's first-level pair sums (segments lane and lane + 32)
__device__ __forceinline__ void final_tree_finish(double s, double q, int lane, uint64_t n_paths, double discount, bool ok,
                                                  unsigned long long epoch, ResultDev *d_out, HostSlot *h_out)
{
    s = warp_fold(s);
    q = warp_fold(q);
    if (lane == 0) {
        MCB_TRACE(5)
        const double n = (double)n_paths;
        const double mean = s / n;
        double var = q / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        if (n_paths > 1) var *= n / (n - 1.0);
        ResultDev r;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        r.price = ok ? discount * mean : nan;
        r.std_error = ok ? discount * sqrt(var / n) : nan;
        r.sum = ok ? s : nan;
        r.sumsq = ok ? q : nan;
        r.n_paths = ok ? n_paths : 0;
        MCB_TRACE(6)
        if (d_out) *d_out = r;
        if (h_out) {
            host_publish(h_out, r, epoch);
            MCB_TRACE(7)
        }
    }
}

__device__ __forceinline__ void final_tree_warp(const volatile double *seg, int lane, uint64_t n_chunks, uint64_t n_paths,
                                                double discount, bool ok, unsigned long long epoch, ResultDev *d_out,
                                                HostSlot *h_out)
{
    const bool lo_live = !segment_is_empty(n_chunks, lane), hi_live = !segment_is_empty(n_chunks, lane + 32);
    const double s = (lo_live ? seg[2 * lane] : 0.0) + (hi_live ? seg[2 * (lane + 32)] : 0.0);
    const double q = (lo_live ? seg[2 * lane + 1] : 0.0) + (hi_live ? seg[2 * (lane + 32) + 1] : 0.0);
    final_tree_finish(s, q, lane, n_paths, discount, ok, epoch, d_out, h_out);
}

// Store segment `seg` = (a, b) of this shard into every consumer's mailbox slot, and -- in the CTA that
// completes the shard's last segment -- publish the flag or (world == 1) run the final tree.
// Called by thread 0's (a, b); every thread of the CTA takes part in the barrier.
__device__ __forceinline__ void job_publish_segment(const JobArgs &args, int seg, double a, double b, int *flag)
{
    const int slot = (int)(args.epoch % (unsigned long long)kRing);
    PeerMailbox *mine = args.peers.box[args.rank];
    if (threadIdx.x == 0) {
        for (int c = 0; c < args.n_consumers; ++c) {
            // the slot still holds its previous sharded job until consumer c has folded it
            if (args.ack_epoch && !wait_at_least(&mine->consumed[c], args.ack_epoch, args.timeout_ns)) {
                atomicAdd(&mine->timeouts, 1u);
                continue;                                   // c will time out on this job and report NaN
            }
            double *dst = args.peers.box[c]->gather[slot] + 2 * seg;
            __stcg(dst, a);
            __stcg(dst + 1, b);
        }
        if (args.world > 1) __threadfence_system(); else __threadfence();
        flag[1] = atomicAdd(&args.seg_tickets[kSegments], 1u) == (unsigned int)args.live_segments - 1u ? 1 : 0;
    }
    __syncthreads();
    MCB_TRACE(4)
    if (!flag[1]) return;
    // ---- the CTA that completed this shard's last segment ------------------------------------
    if (args.world > 1) __threadfence_system(); else __threadfence();   // the flags go to other GPUs: system scope
    if (threadIdx.x == 0) args.seg_tickets[kSegments] = 0u;   // ready for the next launch
    if (args.world > 1) {
        if (threadIdx.x < (unsigned)args.n_consumers)
            *((volatile unsigned long long *)&args.peers.box[threadIdx.x]->flags[slot][args.rank]) = args.epoch;
        return;
    }
    // world == 1: every segment is here; finish the job in this launch
    if (threadIdx.x < 32)
        final_tree_warp(mine->gather[slot], (int)threadIdx.x, args.n_chunks, args.n_paths, args.discount, true,
                        args.epoch, args.d_out, args.h_out);
}

// Fold segment `seg` of this shard (all kSlots threads, same order as segment_kernel) and publish it.
__device__ __forceinline__ void job_fold_and_publish(const JobArgs &args, const float2 *__restrict__ partials,
                                                     uint64_t first_chunk, int seg, double *dscratch, int *flag)
{
    const uint64_t n = args.n_chunks;
    const uint64_t lo = (n * (uint64_t)seg) / kSegments, hi = (n * (uint64_t)(seg + 1)) / kSegments;
    double a = 0.0, b = 0.0;
    for (uint64_t c = lo + threadIdx.x; c < hi; c += kSlots) {
        const float2 v = __ldcg(partials + (c - first_chunk));
        a = a + (double)v.x;
        b = b + (double)v.y;
    }
    block_fold2(a, b, dscratch);
    job_publish_segment(args, seg, a, b, flag);
}

// Tail of european_job_kernel: a ticket per segment finds the last CTA of each.  A segment that
// consists of this one chunk (every job of at most 64 chunks, i.e. the reference's own call sizes)
// needs neither ticket nor fence nor a second look at memory: its 256-slot double tree has a single
// non-zero slot, so the segment IS (double)partial, straight from thread 0's registers.
__device__ __noinline__ void job_tail(const JobArgs &args, const float2 *__restrict__ partials, uint64_t first_chunk,
                                      float2 mine, double *dscratch, int *flag)
{
    const uint64_t chunk = first_chunk + blockIdx.x;
    const uint64_t n = args.n_chunks;
    if (threadIdx.x == 0) {
        int seg = (int)((chunk * (uint64_t)kSegments) / n);
        while ((n * (uint64_t)(seg + 1)) / kSegments <= chunk) ++seg;
        while ((n * (uint64_t)seg) / kSegments > chunk) --seg;
        const uint64_t lo = (n * (uint64_t)seg) / kSegments, hi = (n * (uint64_t)(seg + 1)) / kSegments;
        if (hi - lo == 1) {
            flag[0] = seg | 0x100;                          // alone in its segment
        } else {
            __threadfence();                                // my partial is visible device-wide
            const unsigned int t = atomicAdd(&args.seg_tickets[seg], 1u);
            flag[0] = (t == (unsigned int)(hi - lo) - 1u) ? seg : -1;
            if (flag[0] >= 0) args.seg_tickets[seg] = 0u;   // ready for the next launch
        }
    }
    __syncthreads();
    MCB_TRACE(3)
    const int seg = flag[0];
    if (seg < 0) return;                                    // CTA-uniform: not the last chunk of its segment
    if (seg & 0x100) {
        job_publish_segment(args, seg & 0xff, (double)mine.x, (double)mine.y, flag);
        return;
    }
    __threadfence();
    job_fold_and_publish(args, partials, first_chunk, seg, dscratch, flag);
}

// Large shards: the per-CTA ticket of european_job_kernel holds every CTA's slot for one L2 round
// trip (~0.8 us of 45 us, measured 1.8 % at 2^30 paths), more than a second launch costs once the
// shard runs for ~0.3 ms.  Those jobs price with the plain european_kernel and fold here: one CTA per
// owned segment, then exactly the same publish / final-tree tail.
__global__ void __launch_bounds__(kSlots)
segments_job_kernel(const __grid_constant__ JobArgs args, const float2 *__restrict__ partials, uint64_t first_chunk)
{
    __shared__ double dscratch[2 * kWarps];
    __shared__ int flag[2];
    const int seg = args.seg_lo + (int)blockIdx.x;
    const uint64_t n = args.n_chunks;
    if ((n * (uint64_t)seg) / kSegments == (n * (uint64_t)(seg + 1)) / kSegments) return;   // no chunk: CTA-uniform
    job_fold_and_publish(args, partials, first_chunk, seg, dscratch, flag);
}

template <int TYPE, int PPS>
__global__ void __launch_bounds__(kSlots, 8)
european_job_kernel(const __grid_constant__ EuropeanParams prm, const __grid_constant__ JobArgs args,
                    float2 *__restrict__ partials)
{
    __shared__ float scratch[2 * kWarps];
    __shared__ double dscratch[2 * kWarps];
    __shared__ int flag[2];
    const uint64_t chunk = prm.first_chunk + blockIdx.x;
    const uint64_t base = chunk * (uint64_t)(kSlots * PPS);
    const uint32_t p_hi = (uint32_t)(base >> 32);
    const uint32_t p_lo0 = (uint32_t)base + threadIdx.x;
    const uint64_t left = prm.n_paths - base;
    float sum = 0.0f, sq = 0.0f;
    MCB_TRACE(0)
    if (left >= (uint64_t)(kSlots * PPS)) {
        uint64_t prod1 = (uint64_t)kPhiloxM1 * p_lo0;
#pragma unroll 4
        for (int i = 0; i < PPS; ++i) {
            const float pay = european_payoff_from_prod<TYPE>(prod1, p_hi, prm);
            prod1 += (uint64_t)kPhiloxM1 * kSlots;
            sum = sum + pay;
            sq = fmaf(pay, pay, sq);
        }
    } else {
        for (int i = 0; i < PPS; ++i) {
            const uint32_t local = (uint32_t)(i * kSlots) + threadIdx.x;
            if ((uint64_t)local < left) {
                const float pay = european_payoff<TYPE>(p_lo0 + (uint32_t)(i * kSlots), p_hi, prm);
                sum = sum + pay;
                sq = fmaf(pay, pay, sq);
            }
        }
    }
    MCB_TRACE(1)
    block_fold2(sum, sq, scratch);
    MCB_TRACE(2)
    if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(sum, sq);
    job_tail(args, partials, prm.first_chunk, make_float2(sum, sq), dscratch, flag);
}

// ------------------------------------------------------------------------------------------
// SMALL jobs (at most kSegments chunks = 1 048 576 paths: the reference's own call sizes, hello.cu:14,31 / testing.cu).
// A synchronous call of that size is latency, not throughput: one CTA alone needs 13 750 clk (7 us) for the 64 serial
// path evaluations of its 256 slots (tools/job_latency_probe.cu).  Here a CLUSTER of eight CTAs on eight SMs prices one
// chunk: CTA r evaluates paths [8r, 8r + 8) of every slot (eight independent Philox chains per thread) and sends each
// payoff through distributed shared memory to the CTA that owns the slot -- CTA w sums, in path order, the 32 slots of
// warp w of the chunk, i.e. exactly the additions slot t of european_kernel performs, then warp_fold's five steps --
// and the eight warp totals meet in CTA 0 for block_fold2's last three steps.  Same operands, same order, same bits.
// Both hand-offs are st.async stores that signal an mbarrier in the RECEIVER's shared memory (no fence, no cluster
// barrier on the data path; the one cluster barrier, "everybody has started", hides behind the pricing).  Every chunk
// of such a job is a segment of its own (no segment ticket); the tail is warp-level code in the one warp that is left.
// ------------------------------------------------------------------------------------------
constexpr int kSmallCluster = 8;                 // CTAs (SMs) per chunk
constexpr uint64_t kSmallJobChunks = kSegments;  // jobs of at most this many chunks take this kernel

__device__ __forceinline__ uint32_t cluster_cta_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_map_shared(const void *p, uint32_t rank)   // my smem address as seen in CTA `rank`
{
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
    return out;
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// An mbarrier in MY shared memory that remote st.async stores signal: the receiver learns that the bytes it expects have
// landed without any fence on the sender's side (a barrier.cluster.arrive.release costs a MEMBAR.ALL.GPU: ~0.5 us).
__device__ __forceinline__ void mbar_init_expect(unsigned long long *bar, uint32_t bytes)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar)   // first phase
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], 0;\n"
                     " selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(a)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void cluster_store_signal(uint32_t addr, float v, uint32_t bar)   // both in the same remote CTA
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr),
                 "r"(__float_as_uint(v)), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cluster_store_signal2(uint32_t addr, float v, float w, uint32_t bar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(addr),
                 "r"(__float_as_uint(v)), "r"(__float_as_uint(w)), "r"(bar)
                 : "memory");
}

// Spin (bounded) until the tagged words of this lane's two segments (each only if `want`ed) carry `tag`; the loads of
// both segments are in flight together, so the usual case costs one L2 round trip.  Returns the doubles they spell.
__device__ __forceinline__ bool read_tagged_pair(const unsigned long long *lo, const unsigned long long *hi, bool want_lo,
                                                 bool want_hi, unsigned long long tag, unsigned long long timeout_ns,
                                                 double &lo_a, double &lo_b, double &hi_a, double &hi_b)
{
    unsigned long long t0 = 0;
    while (want_lo || want_hi) {
        unsigned long long x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
        if (want_lo) {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x[0]), "=l"(x[1]) : "l"(lo));
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x[2]), "=l"(x[3]) : "l"(lo + 2));
        }
        if (want_hi) {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(y[0]), "=l"(y[1]) : "l"(hi));
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(y[2]), "=l"(y[3]) : "l"(hi + 2));
        }
        if (want_lo && (x[0] >> 32) == tag && (x[1] >> 32) == tag && (x[2] >> 32) == tag && (x[3] >> 32) == tag) {
            lo_a = __longlong_as_double((long long)((x[0] & 0xffffffffull) | (x[1] << 32)));
            lo_b = __longlong_as_double((long long)((x[2] & 0xffffffffull) | (x[3] << 32)));
            want_lo = false;
        }
        if (want_hi && (y[0] >> 32) == tag && (y[1] >> 32) == tag && (y[2] >> 32) == tag && (y[3] >> 32) == tag) {
            hi_a = __longlong_as_double((long long)((y[0] & 0xffffffffull) | (y[1] << 32)));
            hi_b = __longlong_as_double((long long)((y[2] & 0xffffffffull) | (y[3] << 32)));
            want_hi = false;
        }
        if (want_lo || want_hi) {
            if (t0 == 0) t0 = global_timer_ns();
            else if (global_timer_ns() - t0 > timeout_ns) return false;
        }
    }
    return true;
}

// Warp-level tail of european_small_job_kernel (warp 0 of the cluster's CTA 0; lane 0 holds the chunk's totals), for a
// group of one.  One chunk: the tree runs on registers.  2..64 chunks: every chunk's lane 0 stores
// its segment as four self-tagged 8-byte words and leaves; warp 0 of chunk 0 -- the first cluster to be scheduled,
// so it is waiting while the others still price -- reads the live segments until their tags match (bounded by the
// engine's timeout; the other chunks wait for nobody, so they always get their SMs) and runs the tree.  Neither a
// fence nor an atomic round trip is left on the path from the last payoff to the result (they cost 2 400 + 1 100
// clk, tools/job_latency_probe.cu); the plain copy in `own` serves mcb_last_segments.  Nothing a peer may write is touched,
// so a shard of a larger group can price such a job by itself (args.world == 1, box[0] = its own mailbox).
__device__ __forceinline__ void small_job_tail_one_gpu(const JobArgs &args, float2 *__restrict__ partials, uint64_t chunk,
                                                       float sum, float sq, int lane)
{
    const int slot = (int)(args.epoch % (unsigned long long)kRing);
    PeerMailbox *mine = args.peers.box[0];
    const uint32_t n = (uint32_t)args.n_chunks;
    const double a = (double)sum, b = (double)sq;     // a one-chunk segment IS its partial (see job_tail)
    const unsigned long long tag = args.epoch & 0xffffffffull;
    if (lane == 0) {
        partials[chunk] = make_float2(sum, sq);
        const uint32_t seg = (kSegments * ((uint32_t)chunk + 1u) + n - 1u) / n - 1u;   // n <= kSegments
        double *dst = mine->own[slot] + 2 * seg;
        dst[0] = a;
        dst[1] = b;
        if (n > 1) {
            const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
            volatile unsigned long long *w = mine->tagged[slot][seg];
            w[0] = (tag << 32) | (ua & 0xffffffffull);
            w[1] = (tag << 32) | (ua >> 32);
            w[2] = (tag << 32) | (ub & 0xffffffffull);
            w[3] = (tag << 32) | (ub >> 32);
        }
    }
    MCB_TRACE(4)
    if (n == 1) {
        // the one chunk is segment 63 (lane 31's second operand); every other segment is +0.0 by rule
        const double a0 = __shfl_sync(kFullMask, a, 0), b0 = __shfl_sync(kFullMask, b, 0);
        final_tree_finish(lane == 31 ? 0.0 + a0 : 0.0, lane == 31 ? 0.0 + b0 : 0.0, lane, args.n_paths, args.discount, true,
                          args.epoch, args.d_out, args.h_out);
        return;
    }
    if (chunk != 0) return;
    double lo_a = 0.0, lo_b = 0.0, hi_a = 0.0, hi_b = 0.0;
    bool ok = read_tagged_pair(mine->tagged[slot][lane], mine->tagged[slot][lane + 32], !segment_is_empty(args.n_chunks, lane),
                               !segment_is_empty(args.n_chunks, lane + 32), tag, args.timeout_ns, lo_a, lo_b, hi_a, hi_b);
    __syncwarp();
    if (!ok) atomicAdd(&mine->timeouts, 1u);
    ok = __all_sync(kFullMask, ok);
    final_tree_finish(lo_a + hi_a, lo_b + hi_b, lane, args.n_paths, args.discount, ok, args.epoch, args.d_out, args.h_out);
}

template <int TYPE, int PPS>
__global__ void __cluster_dims__(kSmallCluster, 1, 1) __launch_bounds__(kSlots)
european_small_job_kernel(const __grid_constant__ EuropeanParams prm, const __grid_constant__ JobArgs args,
                          float2 *__restrict__ partials)
{
    static_assert(PPS % kSmallCluster == 0 && kWarps == kSmallCluster, "CTA w of the cluster plays warp w of the chunk");
    constexpr int kPer = PPS / kSmallCluster;            // paths of a slot evaluated by one CTA
    __shared__ float recv[PPS][32];                      // payoffs of MY 32 slots, by path-in-slot
    __shared__ float2 totals[kWarps];                    // CTA 0: the eight warp totals
    __shared__ alignas(8) unsigned long long bars[2];    // [0] my slots' payoffs have landed, [1] (CTA 0) the eight totals have
    const uint32_t rank = cluster_cta_rank();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint64_t chunk = prm.first_chunk + blockIdx.x / kSmallCluster;
    const uint64_t base = chunk * (uint64_t)(kSlots * PPS);
    const uint32_t p_hi = (uint32_t)(base >> 32);
    const uint32_t p_lo0 = (uint32_t)base + (uint32_t)t;
    const uint64_t left64 = prm.n_paths - base;
    const uint32_t left = left64 >= (uint64_t)(kSlots * PPS) ? (uint32_t)(kSlots * PPS) : (uint32_t)left64;
    // paths of slot t that exist: chunk-local t, t + 256, ... < left
    const int cnt = (uint32_t)t < left ? (int)((left - (uint32_t)t + kSlots - 1) / kSlots) : 0;
    // ... and of the slot whose payoffs lane `lane` of warp 0 will sum here (slot 32 rank + lane)
    const uint32_t slot = rank * 32u + (uint32_t)lane;
    const int mine = slot < left ? (int)((left - slot + kSlots - 1) / kSlots) : 0;
    MCB_TRACE(0)
    MCB_TRACE_CTA(0)
    if (warp == 0) {
        const uint32_t expected = __reduce_add_sync(kFullMask, (uint32_t)mine) * (uint32_t)sizeof(float);
        if (lane == 0) {
            mbar_init_expect(&bars[0], expected);
            mbar_init_expect(&bars[1], (uint32_t)(kWarps * sizeof(float2)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    // "every CTA of the cluster has started and its barriers exist" (a CTA's shared memory may only be written by its
    // peers from then on): arrive now, wait just before the first remote store -- the pricing in between hides it
    cluster_arrive_relaxed();
    float pay[kPer] = {};
    if (cnt == PPS) {
#pragma unroll
        for (int j = 0; j < kPer; ++j)
            pay[j] = european_payoff<TYPE>(p_lo0 + (uint32_t)(((int)rank * kPer + j) * kSlots), p_hi, prm);
    } else {
#pragma unroll
        for (int j = 0; j < kPer; ++j)
            if ((int)rank * kPer + j < cnt)
                pay[j] = european_payoff<TYPE>(p_lo0 + (uint32_t)(((int)rank * kPer + j) * kSlots), p_hi, prm);
    }
    __syncwarp();                         // .aligned barriers need all 32 lanes together
    cluster_wait();
    {
        const uint32_t dst = cluster_map_shared(&recv[rank * kPer][lane], (uint32_t)warp);
        const uint32_t bar = cluster_map_shared(&bars[0], (uint32_t)warp);
#pragma unroll
        for (int j = 0; j < kPer; ++j)
            if ((int)rank * kPer + j < cnt) cluster_store_signal(dst + (uint32_t)(j * 32 * sizeof(float)), pay[j], bar);
    }
    MCB_TRACE_CTA(1)
    if (warp != 0) return;
    // ---- warp 0 of CTA `rank` = warp `rank` of the chunk's 256 slots ----
    mbar_wait(&bars[0]);
    MCB_TRACE(1)
    MCB_TRACE_CTA(2)
    float sum = 0.0f, sq = 0.0f;
    if (mine == PPS) {
#pragma unroll 16
        for (int i = 0; i < PPS; ++i) {
            const float pay = recv[i][lane];
            sum = sum + pay;
            sq = fmaf(pay, pay, sq);
        }
    } else {
        for (int i = 0; i < mine; ++i) {
            const float pay = recv[i][lane];
            sum = sum + pay;
            sq = fmaf(pay, pay, sq);
        }
    }
    sum = warp_fold(sum);
    sq = warp_fold(sq);
    if (lane == 0) cluster_store_signal2(cluster_map_shared(&totals[rank], 0u), sum, sq, cluster_map_shared(&bars[1], 0u));
    MCB_TRACE_CTA(3)
    if (rank != 0) return;
    mbar_wait(&bars[1]);
    float x = lane < kWarps ? totals[lane].x : 0.0f, y = lane < kWarps ? totals[lane].y : 0.0f;
#pragma unroll
    for (int off = kWarps / 2; off > 0; off >>= 1) {      // block_fold2's 8 -> 1 step
        x = x + __shfl_down_sync(kFullMask, x, off);
        y = y + __shfl_down_sync(kFullMask, y, off);
    }
    MCB_TRACE(2)
    small_job_tail_one_gpu(args, partials, chunk, x, y, lane);   // (the engine launches this kernel for groups of one only)
}

// A shard that owns no chunk of a (small) job still owes its consumers its flag (its segments are
// empty, i.e. +0.0 by rule).
__global__ void __launch_bounds__(32)
job_publish_empty_kernel(const __grid_constant__ JobArgs args)
{
    const int slot = (int)(args.epoch % (unsigned long long)kRing);
    if (threadIdx.x < (unsigned)args.n_consumers)
        *((volatile unsigned long long *)&args.peers.box[threadIdx.x]->flags[slot][args.rank]) = args.epoch;
}

// Final pass of a world > 1 job on consumer shard `rank`: wait (bounded) until every producer has
// published `epoch` in MY mailbox, run the fixed tree, write the result (device + mapped host),
// then acknowledge the slot to every producer.
__global__ void __launch_bounds__(32)
combine_job_kernel(PeerTable peers, int rank, int world, int spin, unsigned long long epoch,
                   unsigned long long timeout_ns, uint64_t n_chunks, uint64_t n_paths, double discount,
                   ResultDev *__restrict__ d_out, HostSlot *__restrict__ h_out)
{
    const int lane = threadIdx.x;
    const int slot = (int)(epoch % (unsigned long long)kRing);
    PeerMailbox *mine = peers.box[rank];
    bool ok = true;
    if (lane < world) {
        const volatile unsigned long long *f = &mine->flags[slot][lane];
        // spin == 0: the producers' launches were awaited through events (same process), the flag must be there
        ok = spin ? wait_at_least(f, epoch, timeout_ns) : (*f >= epoch);
        if (!ok) atomicAdd(&mine->timeouts, 1u);
    }
    ok = __all_sync(kFullMask, ok);
    __threadfence_system();
    final_tree_warp(mine->gather[slot], lane, n_chunks, n_paths, discount, ok, epoch, d_out, h_out);
    __syncwarp();
    // the slot may be overwritten by job epoch + kRing once every producer sees this ack
    if (lane < world) *((volatile unsigned long long *)&peers.box[lane]->consumed[rank]) = epoch;
}

// Standalone deterministic float sum (reduce3..6 replacement): slot t adds x[t], x[t+256], ...
__global__ void __launch_bounds__(kSlots)
reduce_sum_kernel(const float *__restrict__ x, uint64_t n, float *__restrict__ out)
{
    __shared__ float scratch[2 * kWarps];
    float a = 0.0f, b = 0.0f;
    for (uint64_t i = threadIdx.x; i < n; i += kSlots) a = a + x[i];
    block_fold2(a, b, scratch);
    if (threadIdx.x == 0) out[0] = a;
}

// reduce3..6 index ranges, one CTA per "block" of the reference's launch (see mcb_reduce_blocks).
__global__ void __launch_bounds__(kSlots)
reduce_blocks_kernel(const float *__restrict__ x, uint64_t n, uint64_t span, int strided, float *__restrict__ out)
{
    __shared__ float scratch[2 * kWarps];
    float a = 0.0f, b = 0.0f;
    const uint64_t hop = strided ? (uint64_t)gridDim.x * span : ~0ull;
    for (uint64_t lo = (uint64_t)blockIdx.x * span; lo < n; lo += hop) {
        const uint64_t hi = lo + span < n ? lo + span : n;
        for (uint64_t i = lo + threadIdx.x; i < hi; i += kSlots) a = a + x[i];
        if (!strided) break;
    }
    block_fold2(a, b, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = a;
}

// Pricing from pre-generated normals normals[p*n_steps + i] (inc/trajectories.cuh:14-52).  The
// reference gives a thread a path, so a warp's loads are n_steps*4 bytes apart (32 sectors per
// load); here a WARP takes a path: lane j sums steps j, j+32, ... (coalesced 128-byte loads), the
// lane sums fold 16,8,4,2,1.  Only the terminal price matters, so the walk is one dot product:
//   log2 S_T = log2 S_0 + n_steps*dr + v * sum_i z_i.
__global__ void __launch_bounds__(kSlots)
pregen_kernel(const float *__restrict__ normals, uint64_t n_paths, int n_steps, float l0, float dr, float v,
              float K, float *__restrict__ payoffs)
{
    const int lane = threadIdx.x & 31;
    const uint64_t p = (uint64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (p >= n_paths) return;   // warp-uniform
    const float *z = normals + p * (uint64_t)n_steps;
    float acc = 0.0f;
    if ((n_steps & 3) == 0 && ((uintptr_t)normals & 15) == 0) {   // rows are 16-byte aligned: 128-bit loads
        const float4 *z4 = reinterpret_cast<const float4 *>(z);
        for (int i = lane; i < (n_steps >> 2); i += 32) {
            const float4 q = __ldcs(z4 + i);
            acc = acc + ((q.x + q.y) + (q.z + q.w));
        }
    } else {
        for (int i = lane; i < n_steps; i += 32) acc = acc + __ldcs(z + i);
    }
    acc = warp_fold(acc);
    if (lane == 0) {
        const float l = fmaf(v, acc, fmaf((float)n_steps, dr, l0));
        payoffs[p] = fmaxf(mufu_ex2(l) - K, 0.0f);
    }
}

// ---- parity hooks ---------------------------------------------------------------------
// Exhaustive accuracy scan of the MUFU Box-Muller pieces against double precision over a range
// of 32-bit words: which == 0 -> radius s = sqrt(-2 ln u(x)); 1 -> sin v(y); 2 -> cos v(y).
// out[0] = max |err| (as double bits via atomicMax on the ordered-int image), out[1] = number of
// non-finite or (radius) negative results.
__global__ void __launch_bounds__(256)
boxmuller_scan_kernel(int which, uint64_t first, uint64_t count, unsigned long long *__restrict__ out)
{
    double worst = 0.0;
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(first + i);
        float got;
        double want;
        if (which == 0) {
            got = kSqrt2Ln2 * bm_radius_unscaled(w);
            const double u = (double)fmaf(__uint2float_rn(w), k2Pow32Inv, 0.5f * k2Pow32Inv);
            want = sqrt(-2.0 * log(u));
            if (!(got >= 0.0f)) ++bad;
        } else {
            const float v = bm_angle(w);
            got = which == 1 ? mufu_sin(v) : mufu_cos(v);
            // the oracle's angle: cuRAND's unsigned map, the same angle modulo 2 pi
            const double vu = (double)fmaf(__uint2float_rn(w), k2Pow32Inv2Pi, 0.5f * k2Pow32Inv2Pi);
            want = which == 1 ? sin(vu) : cos(vu);
        }
        if (!isfinite(got)) ++bad;
        const double err = fabs((double)got - want);
        worst = err > worst ? err : worst;
    }
    // non-negative doubles order like their bit patterns
    atomicMax(out, (unsigned long long)__double_as_longlong(worst));
    if (bad) atomicAdd(out + 1, bad);
}


__global__ void philox_blocks_kernel(PhiloxKeys keys, const uint64_t *__restrict__ subseq,
                                     const uint64_t *__restrict__ block, uint64_t n, uint4 *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = subseq[i], b = block[i];
    const Words4 w = philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)s, (uint32_t)(s >> 32), keys);
    out[i] = make_uint4(w.x, w.y, w.z, w.w);
}

__global__ void stream_normals_kernel(PhiloxKeys keys, uint64_t subseq, uint64_t n0, uint64_t count,
                                      float *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t n = n0 + i, b = n >> 2;
    float z[4];
    increments4(philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32), keys),
                kSqrt2Ln2, 0.0f, z);
    out[i] = z[n & 3];
}

}  // namespace mcb
