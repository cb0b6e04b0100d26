/*
 * mc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see mc_oracle.h for the parity status).
 *
 * Every function cites the reference (amauryrlm/Monte-Carlo-Project-CUDA, paths relative
 * to /root/reference) or the third-party header (cuRAND 10.3.10, CUDA 12.9,
 * /usr/local/cuda/include) whose arithmetic it restates.  Nothing here is copied: the
 * reference draws XORWOW normals from per-thread state; this file draws the SAME
 * SHAPE of stream -- curand_init(seed, subsequence = path id, offset = 0), inc/tool.cuh:194 --
 * from the stateless Philox4x32-10 generator and does the GBM maths in double.
 */
#include "mc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10.  cuRAND: curand_philox4x32_x.h:88-91 (constants), :160-170 (round),
 * :172-192 (ten rounds with key bump between rounds).
 * ---------------------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout of curand_init(seed, subsequence, 0, curandStatePhilox4_32_10_t*):
 * key = (seed_lo, seed_hi) curand_kernel.h:1028-1029; skipahead_sequence adds the
 * subsequence to ctr.zw (curand_philox4x32_x.h:123-133); every curand4() bumps
 * ctr.xy by one (curand_philox4x32_x.h:137-143).  Block b of path p is therefore
 * Philox(ctr = (b_lo, b_hi, p_lo, p_hi)).  The reference call site this mirrors is
 * curand_init(seed, tid, 0, &state[tid]) at inc/tool.cuh:194. */
void orc_stream_block(uint64_t seed, uint64_t subsequence, uint64_t block, uint32_t out[4])
{
    uint32_t ctr[4] = { (uint32_t)block, (uint32_t)(block >> 32),
                        (uint32_t)subsequence, (uint32_t)(subsequence >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
}

/* _curand_box_muller, curand_normal.h:70-88 with curand_globals.h:56-60 constants.
 * u in (0,1], v in (0, 2pi]; fmaf mirrors the FFMA the device emits. */
float orc_uniform_u(uint32_t x)
{
    return fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

float orc_angle_v(uint32_t y)
{
    const float k = 2.3283064e-10f * 6.2831855f;
    return fmaf((float)y, k, k / 2.0f);
}

/* curand_normal(curandStatePhilox4_32_10_t*), curand_normal.h:345-360: normals come in
 * pairs from consecutive words (x, y): even index -> s*sin(v), odd index -> s*cos(v). */
static void box_muller_pair(uint32_t x, uint32_t y, double *n_sin, double *n_cos)
{
    double u = (double)orc_uniform_u(x);
    double v = (double)orc_angle_v(y);
    double s = sqrt(-2.0 * log(u));
    *n_sin = s * sin(v);
    *n_cos = s * cos(v);
}

double orc_stream_normal(uint64_t seed, uint64_t subsequence, uint64_t n)
{
    uint32_t w[4];
    double a, b;
    orc_stream_block(seed, subsequence, n >> 2, w);
    int pair = (int)((n >> 1) & 1);
    box_muller_pair(w[2 * pair], w[2 * pair + 1], &a, &b);
    return (n & 1) ? b : a;
}

void orc_stream_normals(uint64_t seed, uint64_t subsequence, uint64_t n0, uint64_t count, double *out)
{
    uint32_t w[4];
    double z[4];
    uint64_t cached = UINT64_MAX;
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t n = n0 + i;
        if ((n >> 2) != cached) {
            cached = n >> 2;
            orc_stream_block(seed, subsequence, cached, w);
            box_muller_pair(w[0], w[1], &z[0], &z[1]);
            box_muller_pair(w[2], w[3], &z[2], &z[3]);
        }
        out[i] = z[n & 3];
    }
}

/* ------------------------------------------------------------------------------------------
 * European option, one step.  Restates simulateOptionPriceMultipleBlockGPUwithReduce
 * (inc/trajectories.cuh:71-76) and its CPU twin simulateOptionPriceCPU (inc/tool.cuh:119-126):
 *   St = S0 * exp((r - sigma^2/2) T + sigma sqrt(T) G);  payoff = max(St - K, 0).
 * Path p uses normal 0 of subsequence p.  The put payoff max(K - St, 0) is the additive
 * capability north_star asks for (no counterpart in the reference).
 * ---------------------------------------------------------------------------------------- */
static double payoff_of(double St, double K, int option_type)
{
    double x = (option_type == ORC_PUT) ? (K - St) : (St - K);
    return x > 0.0 ? x : 0.0;
}

void orc_european(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                  int option_type, double *sum, double *sumsq, float *payoffs)
{
    double S0 = o->S0, K = o->K, r = o->r, sig = o->v, T = o->T;
    double drift = (r - 0.5 * sig * sig) * T;
    double vol = sig * sqrt(T);
    double s = 0.0, s2 = 0.0;
    for (uint64_t i = 0; i < n_paths; ++i) {
        double G = orc_stream_normal(seed, first_path + i, 0);
        double St = S0 * exp(drift + vol * G);
        double p = payoff_of(St, K, option_type);
        s += p;
        s2 += p * p;
        if (payoffs) payoffs[i] = (float)p;
    }
    if (sum) *sum = s;
    if (sumsq) *sumsq = s2;
}

/* The same pricer under PACKED keying ("European path, packed keying"): path p draws normal
 * p & 3 of the cuRAND stream (seed, subsequence p >> 2) -- what four successive curand_normal() calls on one
 * state return (curand_normal.h:345-360), the reference's own way of consuming a generator
 * (inc/trajectories.cuh:145) -- instead of normal 0 of subsequence p.  All four normals of a Philox block are
 * used, so a path costs a quarter of a block. */
void orc_european_packed(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                         int option_type, double *sum, double *sumsq, float *payoffs)
{
    double S0 = o->S0, K = o->K, r = o->r, sig = o->v, T = o->T;
    double drift = (r - 0.5 * sig * sig) * T;
    double vol = sig * sqrt(T);
    double s = 0.0, s2 = 0.0;
    for (uint64_t i = 0; i < n_paths; ++i) {
        uint64_t path = first_path + i;
        double G = orc_stream_normal(seed, path >> 2, path & 3);
        double St = S0 * exp(drift + vol * G);
        double p = payoff_of(St, K, option_type);
        s += p;
        s2 += p * p;
        if (payoffs) payoffs[i] = (float)p;
    }
    if (sum) *sum = s;
    if (sumsq) *sumsq = s2;
}

/* ------------------------------------------------------------------------------------------
 * Bullet (barrier-count) option.  Restates simulateBulletOptionPriceMultipleBlockGPU
 * (inc/trajectories.cuh:138-153) and simulateBulletOptionPriceCPU (inc/tool.cuh:155-171):
 * N_STEPS - Tk steps of dt = option.step; count += (B > St) each step; payoff only if
 * P1 <= count <= P2; optional restart state (Ik, Sk, Tk) with Sk == 0 meaning "start at S0".
 * The walk is done in log space (log St < log B), which is what the engine does.
 * ---------------------------------------------------------------------------------------- */
void orc_bullet(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                int Ik, float Sk, int Tk, double *sum, double *sumsq, float *payoffs)
{
    double K = o->K, r = o->r, sig = o->v, dt = o->step;
    double drift = (r - 0.5 * sig * sig) * dt;
    double vol = sig * sqrt(dt);
    double logB = (o->B > 0.0f) ? log((double)o->B) : -INFINITY;
    double start = (Sk == 0.0f) ? (double)o->S0 : (double)Sk;
    int steps = o->N_STEPS - Tk;
    double *z = (double *)malloc(sizeof(double) * (size_t)(steps > 0 ? steps : 1));
    double s = 0.0, s2 = 0.0;
    for (uint64_t i = 0; i < n_paths; ++i) {
        double logS = log(start);
        int count = Ik;
        if (steps > 0) orc_stream_normals(seed, first_path + i, 0, (uint64_t)steps, z);
        for (int k = 0; k < steps; ++k) {
            logS += drift + vol * z[k];
            if (logS < logB) count += 1;
        }
        double p = 0.0;
        if (count >= o->P1 && count <= o->P2) p = payoff_of(exp(logS), K, ORC_CALL);
        s += p;
        s2 += p * p;
        if (payoffs) payoffs[i] = (float)p;
    }
    free(z);
    if (sum) *sum = s;
    if (sumsq) *sumsq = s2;
}

/* ------------------------------------------------------------------------------------------
 * Full trajectories, path-major: prices[p*N_STEPS + i] = S(t_{i+1}), counts[...] = barrier
 * count after step i+1.  Restates simulate_outer_trajectories (inc/trajectories.cuh:296-306)
 * and kernel B of Simulation::simulate_outer_trajectories (inc/testing.cuh:62-72).
 * S0 itself is not stored (the CSV writer injects it, testing.cu:44).
 * ---------------------------------------------------------------------------------------- */
void orc_trajectories(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                      float *prices, int *counts)
{
    double r = o->r, sig = o->v, dt = o->step;
    double drift = (r - 0.5 * sig * sig) * dt;
    double vol = sig * sqrt(dt);
    double logB = (o->B > 0.0f) ? log((double)o->B) : -INFINITY;
    int steps = o->N_STEPS;
    double *z = (double *)malloc(sizeof(double) * (size_t)(steps > 0 ? steps : 1));
    for (uint64_t i = 0; i < n_paths; ++i) {
        double logS = log((double)o->S0);
        int count = 0;
        orc_stream_normals(seed, first_path + i, 0, (uint64_t)steps, z);
        for (int k = 0; k < steps; ++k) {
            logS += drift + vol * z[k];
            if (logS < logB) count += 1;
            prices[i * (uint64_t)steps + k] = (float)exp(logS);
            if (counts) counts[i * (uint64_t)steps + k] = count;
        }
    }
    free(z);
}

/* ------------------------------------------------------------------------------------------
 * Nested Monte Carlo.  Intent of compute_nmc_one_block_per_point (inc/nmc.cuh:47-104):
 * for every outer point (path p, step k) with state (S, I) = (prices[p,k], counts[p,k]):
 *   remaining = N_STEPS - (k+1)                                  (inc/nmc.cuh:48)
 *   skip (F = 0) when I > P2                                     (inc/nmc.cuh:53)
 *   every inner path restarts from (S, I)  -- the reference forgets to (inc/nmc.cuh:51-53);
 *   payoff gated by P1 <= I <= P2                                (inc/nmc.cuh:60-61)
 *   F = discount * mean over N_PATHS_INNER                       (inc/nmc.cuh:101)
 * discount = exp(-rT) in COMPAT mode (what the reference does), exp(-r (T - t_{k+1})) in
 * CORRECT mode.  Outer paths: seed_outer, subsequence p (wrappers use 1234,
 * inc/wrappers.cuh:151).  Inner path j of point q = p*N_STEPS + k: seed_inner (1235,
 * inc/wrappers.cuh:163), subsequence q*N_PATHS_INNER + j, normal i for inner step i.
 * ---------------------------------------------------------------------------------------- */
void orc_nmc(const orc_option_data *o, uint64_t first_outer, uint64_t n_outer, uint64_t seed_outer,
             uint64_t seed_inner, int discount_mode, float *F, float *prices, int *counts)
{
    int steps = o->N_STEPS, n_inner = o->N_PATHS_INNER;
    double K = o->K, r = o->r, sig = o->v, dt = o->step, T = o->T;
    double drift = (r - 0.5 * sig * sig) * dt;
    double vol = sig * sqrt(dt);
    double logB = (o->B > 0.0f) ? log((double)o->B) : -INFINITY;
    float *pr = (float *)malloc(sizeof(float) * (size_t)steps);
    int *ct = (int *)malloc(sizeof(int) * (size_t)steps);
    double *zo = (double *)malloc(sizeof(double) * (size_t)steps);
    double *z = (double *)malloc(sizeof(double) * (size_t)steps);
    for (uint64_t a = 0; a < n_outer; ++a) {
        uint64_t p = first_outer + a;
        /* outer walk: keep the log-price in double so the inner start is the exact state */
        double logS = log((double)o->S0);
        int count = 0;
        orc_stream_normals(seed_outer, p, 0, (uint64_t)steps, zo);
        for (int k = 0; k < steps; ++k) {
            logS += drift + vol * zo[k];
            if (logS < logB) count += 1;
            pr[k] = (float)exp(logS);
            ct[k] = count;
            if (prices) prices[a * (uint64_t)steps + k] = pr[k];
            if (counts) counts[a * (uint64_t)steps + k] = count;

            int remaining = steps - (k + 1);
            double acc = 0.0;
            if (count <= o->P2) {
                uint64_t q = p * (uint64_t)steps + (uint64_t)k;
                for (int j = 0; j < n_inner; ++j) {
                    double ls = logS;
                    int c = count;
                    if (remaining > 0)
                        orc_stream_normals(seed_inner, q * (uint64_t)n_inner + (uint64_t)j, 0,
                                           (uint64_t)remaining, z);
                    for (int i = 0; i < remaining; ++i) {
                        ls += drift + vol * z[i];
                        if (ls < logB) c += 1;
                    }
                    if (c >= o->P1 && c <= o->P2) acc += payoff_of(exp(ls), K, ORC_CALL);
                }
            }
            double disc = (discount_mode == ORC_DISCOUNT_CORRECT)
                              ? exp(-r * (T - (double)(k + 1) * dt))
                              : exp(-r * T);
            F[a * (uint64_t)steps + k] = (float)(disc * acc / (double)n_inner);
        }
    }
    free(pr); free(ct); free(zo); free(z);
}

/* ------------------------------------------------------------------------------------------
 * Batched strike/vol sweep with common random numbers (BASELINE config 5): parameter set i
 * is priced on exactly the stream a separate European call with (K_i, sigma_i) would see.
 * ---------------------------------------------------------------------------------------- */
void orc_sweep(const orc_option_data *o, const float *strikes, const float *vols, int n_params,
               uint64_t first_path, uint64_t n_paths, uint64_t seed, int option_type,
               double *sums, double *sumsqs)
{
    double S0 = o->S0, r = o->r, T = o->T;
    for (int i = 0; i < n_params; ++i) { sums[i] = 0.0; sumsqs[i] = 0.0; }
    for (uint64_t a = 0; a < n_paths; ++a) {
        double G = orc_stream_normal(seed, first_path + a, 0);
        for (int i = 0; i < n_params; ++i) {
            double sig = vols[i];
            double St = S0 * exp((r - 0.5 * sig * sig) * T + sig * sqrt(T) * G);
            double p = payoff_of(St, (double)strikes[i], option_type);
            sums[i] += p;
            sumsqs[i] += p * p;
        }
    }
}

/* Pricing from pre-generated normals, normals[p*n_steps + i].  Restates
 * simulateOptionPriceMultipleBlockGPU (inc/trajectories.cuh:38-52) and the CPU overload
 * simulateOptionPriceCPU(..., h_randomData, ...) (inc/testing.cuh:75-91). */
void orc_pregen_european(const orc_option_data *o, const float *normals, uint64_t n_paths, int n_steps,
                         float *payoffs)
{
    double r = o->r, sig = o->v, dt = o->step;
    double drift = (r - 0.5 * sig * sig) * dt, vol = sig * sqrt(dt);
    for (uint64_t p = 0; p < n_paths; ++p) {
        double logS = log((double)o->S0);
        for (int i = 0; i < n_steps; ++i) logS += drift + vol * (double)normals[p * (uint64_t)n_steps + i];
        payoffs[p] = (float)payoff_of(exp(logS), (double)o->K, ORC_CALL);
    }
}

/* ------------------------------------------------------------------------------------------
 * Host finalise: price = exp(-rT) * sum / N  (inc/wrappers.cuh:51,85,118).  The reference has
 * no standard error (reduce.cuh sums only); SE follows from the sum of squares.
 * ---------------------------------------------------------------------------------------- */
double orc_price_from_sum(double sum, uint64_t n_paths, float r, float T)
{
    return exp(-(double)r * (double)T) * sum / (double)n_paths;
}

double orc_std_error(double sum, double sumsq, uint64_t n_paths, float r, float T)
{
    double n = (double)n_paths;
    double mean = sum / n;
    double var = sumsq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    if (n_paths > 1) var *= n / (n - 1.0);
    return exp(-(double)r * (double)T) * sqrt(var / n);
}

/* CND: Abramowitz-Stegun 26.2.17 five-term polynomial in float, inc/BlackandScholes.hpp:8-30. */
float orc_cnd_reference(float x)
{
    const float p = 0.2316419f;
    const float b[5] = { 0.31938153f, -0.356563782f, 1.781477937f, -1.821255978f, 1.330274429f };
    const float inv_sqrt_2pi = 0.39894228f;
    float ax = x >= 0.0f ? x : -x;
    float t = 1.0f / (1.0f + p * ax);
    float poly = t * (t * (t * (t * b[4] + b[3]) + b[2]) + b[1]) + b[0];
    float tail = inv_sqrt_2pi * expf(-x * x / 2.0f) * t * poly;
    return x >= 0.0f ? 1.0f - tail : tail;
}

/* black_scholes_CPU, inc/BlackandScholes.hpp:34-43.  The reference mixes precisions: the
 * 0.5 literal promotes the d1 numerator to double, while exp(-rT) is C++'s float overload
 * std::exp(float) (the argument is a float and <cmath> is in scope), i.e. expf. */
float orc_bs_call_reference(float S0, float K, float T, float r, float v)
{
    float sqrtT = sqrtf(T);
    float d1 = (float)((logf(S0 / K) + (r + 0.5 * v * v) * T) / (v * sqrtT));
    float d2 = d1 - v * sqrtT;
    float n1 = orc_cnd_reference(d1);
    float n2 = orc_cnd_reference(d2);
    return S0 * n1 - K * expf(-r * T) * n2;
}

static double phi_exact(double x) { return 0.5 * erfc(-x / sqrt(2.0)); }

double orc_bs_call_exact(double S0, double K, double T, double r, double v)
{
    double d1 = (log(S0 / K) + (r + 0.5 * v * v) * T) / (v * sqrt(T));
    double d2 = d1 - v * sqrt(T);
    return S0 * phi_exact(d1) - K * exp(-r * T) * phi_exact(d2);
}

double orc_bs_put_exact(double S0, double K, double T, double r, double v)
{
    return orc_bs_call_exact(S0, K, T, r, v) - S0 + K * exp(-r * T);
}

/* ------------------------------------------------------------------------------------------
 * The engine's deterministic reduction (replaces reduce3..6, inc/reduce.cuh:9-227, and the
 * tree inlined at inc/trajectories.cuh:77-111, float sum only + atomicAdd).  Restated in
 * the exact operation order so tests can demand BIT-identical partials for given payoffs:
 *   slot t of 256 accumulates chunk-local paths t, t+256, t+512, ... in that order
 *   (sum: +, sum of squares: fma);  lanes fold 16,8,4,2,1;  the 8 warp sums fold 4,2,1.
 * ---------------------------------------------------------------------------------------- */
static float tree256_f32(float *x)
{
    for (int w = 0; w < 8; ++w)
        for (int off = 16; off > 0; off >>= 1)
            for (int j = 0; j < off; ++j) x[w * 32 + j] = x[w * 32 + j] + x[w * 32 + j + off];
    float y[8];
    for (int w = 0; w < 8; ++w) y[w] = x[w * 32];
    for (int off = 4; off > 0; off >>= 1)
        for (int j = 0; j < off; ++j) y[j] = y[j] + y[j + off];
    return y[0];
}

static double tree256_f64(double *x)
{
    for (int w = 0; w < 8; ++w)
        for (int off = 16; off > 0; off >>= 1)
            for (int j = 0; j < off; ++j) x[w * 32 + j] = x[w * 32 + j] + x[w * 32 + j + off];
    double y[8];
    for (int w = 0; w < 8; ++w) y[w] = x[w * 32];
    for (int off = 4; off > 0; off >>= 1)
        for (int j = 0; j < off; ++j) y[j] = y[j] + y[j + off];
    return y[0];
}

void orc_chunk_tree_f32(const float *payoffs, uint64_t n_valid, int paths_per_slot, float *sum, float *sumsq)
{
    float s[ORC_SLOTS], q[ORC_SLOTS];
    for (int t = 0; t < ORC_SLOTS; ++t) {
        float a = 0.0f, b = 0.0f;
        for (int i = 0; i < paths_per_slot; ++i) {
            uint64_t idx = (uint64_t)i * ORC_SLOTS + (uint64_t)t;
            if (idx < n_valid) {
                float p = payoffs[idx];
                a = a + p;
                b = fmaf(p, p, b);
            }
        }
        s[t] = a;
        q[t] = b;
    }
    *sum = tree256_f32(s);
    *sumsq = tree256_f32(q);
}

void orc_segment_range(uint64_t n_chunks, int segment, uint64_t *lo, uint64_t *hi)
{
    *lo = (n_chunks * (uint64_t)segment) / ORC_SEGMENTS;
    *hi = (n_chunks * (uint64_t)(segment + 1)) / ORC_SEGMENTS;
}

void orc_segment_tree_f64(const float *chunk_partials, uint64_t n_chunks, double *segments)
{
    for (int sgm = 0; sgm < ORC_SEGMENTS; ++sgm) {
        uint64_t lo, hi;
        double s[ORC_SLOTS], q[ORC_SLOTS];
        orc_segment_range(n_chunks, sgm, &lo, &hi);
        for (int t = 0; t < ORC_SLOTS; ++t) {
            double a = 0.0, b = 0.0;
            for (uint64_t c = lo + (uint64_t)t; c < hi; c += ORC_SLOTS) {
                a = a + (double)chunk_partials[2 * c];
                b = b + (double)chunk_partials[2 * c + 1];
            }
            s[t] = a;
            q[t] = b;
        }
        segments[2 * sgm] = tree256_f64(s);
        segments[2 * sgm + 1] = tree256_f64(q);
    }
}

void orc_final_tree_f64(const double *segments, double *sum, double *sumsq)
{
    double s[32], q[32];
    for (int j = 0; j < 32; ++j) {
        s[j] = segments[2 * j] + segments[2 * (j + 32)];
        q[j] = segments[2 * j + 1] + segments[2 * (j + 32) + 1];
    }
    for (int off = 16; off > 0; off >>= 1)
        for (int j = 0; j < off; ++j) {
            s[j] = s[j] + s[j + off];
            q[j] = q[j] + q[j + off];
        }
    *sum = s[0];
    *sumsq = q[0];
}

/* Plain float sum of an array through the same slot/tree order with a single chunk
 * (the standalone reduce3..6 replacement, inc/reduce.cuh). */
float orc_reduce_sum_f32(const float *x, uint64_t n)
{
    float s[ORC_SLOTS];
    for (int t = 0; t < ORC_SLOTS; ++t) {
        float a = 0.0f;
        for (uint64_t i = (uint64_t)t; i < n; i += ORC_SLOTS) a = a + x[i];
        s[t] = a;
    }
    return tree256_f32(s);
}
