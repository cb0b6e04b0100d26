/*
 * mc_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's GBM Monte Carlo hot path
 * (amauryrlm/Monte-Carlo-Project-CUDA), re-keyed on the counter-based
 * Philox4x32-10 stream the B200 engine uses.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity status
 *   - integer RNG stream + path indexing : PINNED (Random123 KATs + cuRAND 10.3.10
 *     Philox header compiled for host, tests/golden/philox_vectors.json)
 *   - closed form                        : PINNED against the reference's own
 *     black_scholes_CPU compiled from /root/reference (tests/golden/reference_cpu.json)
 *   - Monte Carlo prices                 : the reference ships no golden vectors and
 *     seeds its CPU path from std::random_device (inc/tool.cuh:116,151), so prices
 *     are pinned STATISTICALLY (3 SE) against oracle/_ref and the closed form.
 *   - nested Monte Carlo                 : PARITY UNPINNED by the reference (three
 *     mutually inconsistent kernels, no CPU twin); this file restates the intent
 *     of inc/nmc.cuh:47-104 and is checked by closed-form limits.
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 48-byte layout as the reference's `struct OptionData` (inc/tool.cuh:13-26). */
typedef struct {
    float S0, T, K, r, v, B;
    int P1, P2, N_PATHS, N_PATHS_INNER, N_STEPS;
    float step;
} orc_option_data;

enum { ORC_CALL = 0, ORC_PUT = 1 };
enum { ORC_DISCOUNT_COMPAT = 0, ORC_DISCOUNT_CORRECT = 1 };

/* Reduction-tree geometry shared with the engine (include/mcb200.h). */
enum {
    ORC_SLOTS = 256,          /* accumulation slots (= CTA threads) per chunk        */
    ORC_SEGMENTS = 64         /* double-precision segments the chunk range is cut in */
};

/* ---- RNG: Philox4x32-10 exactly as cuRAND implements it ------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_stream_block(uint64_t seed, uint64_t subsequence, uint64_t block, uint32_t out[4]);
float orc_uniform_u(uint32_t x);
float orc_angle_v(uint32_t y);
double orc_stream_normal(uint64_t seed, uint64_t subsequence, uint64_t n);
void orc_stream_normals(uint64_t seed, uint64_t subsequence, uint64_t n0, uint64_t count, double *out);

/* ---- pricers (double maths on the float uniforms) -------------------------------- */
void orc_european(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                  int option_type, double *sum, double *sumsq, float *payoffs /* nullable */);
void orc_european_packed(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                  int option_type, double *sum, double *sumsq, float *payoffs /* nullable */);
void orc_bullet(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                int Ik, float Sk, int Tk, double *sum, double *sumsq, float *payoffs /* nullable */);
void orc_trajectories(const orc_option_data *o, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                      float *prices, int *counts /* nullable */);
void orc_nmc(const orc_option_data *o, uint64_t first_outer, uint64_t n_outer, uint64_t seed_outer,
             uint64_t seed_inner, int discount_mode, float *F, float *prices /* nullable */,
             int *counts /* nullable */);
void orc_sweep(const orc_option_data *o, const float *strikes, const float *vols, int n_params,
               uint64_t first_path, uint64_t n_paths, uint64_t seed, int option_type,
               double *sums, double *sumsqs);
void orc_pregen_european(const orc_option_data *o, const float *normals, uint64_t n_paths, int n_steps,
                         float *payoffs);

/* ---- host finalise + closed forms -------------------------------------------------- */
double orc_price_from_sum(double sum, uint64_t n_paths, float r, float T);
double orc_std_error(double sum, double sumsq, uint64_t n_paths, float r, float T);
float orc_cnd_reference(float x);
float orc_bs_call_reference(float S0, float K, float T, float r, float v);
double orc_bs_call_exact(double S0, double K, double T, double r, double v);
double orc_bs_put_exact(double S0, double K, double T, double r, double v);

/* ---- the engine's deterministic reduction tree, restated bit-for-bit in C --------- */
void orc_chunk_tree_f32(const float *payoffs, uint64_t n_valid, int paths_per_slot, float *sum, float *sumsq);
void orc_segment_range(uint64_t n_chunks, int segment, uint64_t *lo, uint64_t *hi);
void orc_segment_tree_f64(const float *chunk_partials /* [n_chunks][2] */, uint64_t n_chunks,
                          double *segments /* [ORC_SEGMENTS][2] */);
void orc_final_tree_f64(const double *segments /* [ORC_SEGMENTS][2] */, double *sum, double *sumsq);
float orc_reduce_sum_f32(const float *x, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
