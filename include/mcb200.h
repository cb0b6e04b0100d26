/*
 * mcb200.h -- C-ABI of the B200-native Monte Carlo pricing engine (libmcb200.so).
 *
 * This is the drop-in boundary for the GBM hot path of amauryrlm/Monte-Carlo-Project-CUDA.
 * The reference has no FFI layer: its boundary is the set of free functions in
 * inc/wrappers.cuh plus Simulation::simulate_outer_trajectories (inc/testing.cuh:281).
 * Every entry point below names the reference interface it replaces (file:line relative
 * to the reference repo).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns an mcb_status; mcb_last_error() gives the message.  The
 *     library never prints and never calls exit() (the reference does both:
 *     inc/tool.cuh:92-100, inc/wrappers.cuh:52).
 *   - there is NO CPU fallback: without a CUDA device mcb_engine_create fails.
 *   - RNG: stateless Philox4x32-10.  Path p of a run with seed s draws from the cuRAND
 *     stream curand_init(s, subsequence = p, offset = 0, curandStatePhilox4_32_10_t*) --
 *     the same call shape as the reference's curand_init(seed, tid, 0, ...) at
 *     inc/tool.cuh:194.  Step i of a path uses normal i of that stream
 *     (curand_normal order: even -> s*sin, odd -> s*cos, curand_normal.h:345-360).
 *   - results are a pure function of (parameters, seed, n_paths): independent of grid
 *     shape, of threadsPerBlock hints and of the number of GPUs.
 */
#ifndef MCB200_H
#define MCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_VERSION 2

/* Geometry of the deterministic reduction (replaces inc/reduce.cuh + the float atomics). */
#define MCB_SLOTS 256                   /* accumulation slots per chunk = CTA threads          */
#define MCB_SEGMENTS 64                 /* double-precision segments per run                  */
#define MCB_EUROPEAN_PATHS_PER_SLOT 64  /* European / sweep chunk = 16384 paths               */
#define MCB_BULLET_PATHS_PER_SLOT 4     /* bullet chunk = 1024 paths                          */

/* Byte-compatible with the reference's `struct OptionData` (inc/tool.cuh:13-26, 48 bytes). */
typedef struct mcb_option_data {
    float S0, T, K, r, v, B;
    int P1, P2, N_PATHS, N_PATHS_INNER, N_STEPS;
    float step;
} mcb_option_data;

/* Replaces the bare `float` the wrappers return (inc/wrappers.cuh:51-56). */
typedef struct mcb_result {
    double price;      /* exp(-rT) * sum / n_paths                       (inc/wrappers.cuh:51) */
    double std_error;  /* from the sum of squares; the reference has none                      */
    double sum;        /* undiscounted payoff sum                                              */
    double sumsq;      /* undiscounted sum of squared payoffs                                  */
    uint64_t n_paths;
} mcb_result;

typedef struct mcb_device_info {
    char name[128];
    int sm_count;
    int cc_major, cc_minor;
    int clock_khz;
    size_t total_mem;
} mcb_device_info;

typedef enum mcb_status {
    MCB_OK = 0,
    MCB_ERR_INVALID = 1,   /* bad argument                                  */
    MCB_ERR_CUDA = 2,      /* CUDA runtime error (see mcb_last_error)       */
    MCB_ERR_NO_DEVICE = 3, /* no usable sm_100 device                       */
    MCB_ERR_NOMEM = 4,
    MCB_ERR_TIMEOUT = 5    /* a peer shard never delivered (result poisoned with NaN) */
} mcb_status;

enum { MCB_CALL = 0, MCB_PUT = 1 };
enum { MCB_DISCOUNT_COMPAT = 0,   /* exp(-rT) for every k, as inc/nmc.cuh:101,268,379 */
       MCB_DISCOUNT_CORRECT = 1 };/* exp(-r (T - t_{k+1}))                             */
enum { MCB_HOST = 0, MCB_DEVICE = 1 }; /* where a caller-provided buffer lives */

typedef struct mcb_engine mcb_engine;

/* ---- engine lifetime.  Replaces the per-call cudaMalloc/cudaFree of every wrapper
 *      (inc/wrappers.cuh:38-55) with a persistent handle owning stream + workspaces. ---- */
int mcb_engine_create(int device, mcb_engine **out);
/* ONE engine over several GPUs of this process (SURVEY.md 8(b) "multi-GPU handled inside the
 * engine", 8(e)): devices[0] leads.  Every whole-job call below then shards internally by path
 * index -- European / bullet / sweep by reduction segments (the (sum, sumsq) segments cross NVLink as
 * plain peer stores from inside the kernels; no collective library), trajectories and nested MC by
 * contiguous path slabs (host destinations) -- and returns the SAME bits as a single-device engine.
 * The reference is single-GPU (inc/wrappers.cuh:33-57 runs on the implicit device 0); a caller of
 * wrapper_gpu_option_vanilla gets all GPUs by creating its engine here.  A device may be listed
 * more than once (several shards on one GPU: used by the single-GPU tests of the sharded path).
 * Buffers passed with MCB_DEVICE must live on devices[0]; such calls run on the leader alone. */
int mcb_engine_create_multi(const int *devices, int n_devices, mcb_engine **out);
int mcb_engine_shard_count(mcb_engine *e);
int mcb_engine_destroy(mcb_engine *e);
const char *mcb_last_error(void);
int mcb_version(void);
int mcb_get_device_info(mcb_engine *e, mcb_device_info *out);   /* getDeviceProperty, inc/tool.cuh:56-88 */
int mcb_synchronize(mcb_engine *e);

/* ---- whole-job, synchronous entry points ------------------------------------------- */

/* European call/put, single step.  Replaces wrapper_gpu_option_vanilla
 * (inc/wrappers.cuh:33-57) = setup_kernel (inc/tool.cuh:192) +
 * simulateOptionPriceMultipleBlockGPUwithReduce (inc/trajectories.cuh:54-113) + host
 * finalise.  n_paths == 0 means opt->N_PATHS.  Reference defaults: seed 1234, MCB_CALL. */
int mcb_price_european(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                       int option_type, mcb_result *out);

/* The same estimator under PACKED keying (no reference counterpart; SURVEY.md 8(d) lists it as the optional
 * second keying): path p draws normal p & 3 of the stream (seed, subsequence p >> 2) -- what four successive
 * curand_normal() calls on one state return -- so one Philox block prices four paths (about 2.5x the paths/s of the
 * canonical keying above, which stays the default and the headline).  A different, equally valid assignment of
 * random numbers to paths: prices agree with mcb_price_european statistically, not bit for bit.  Works on
 * multi-device engines; results do not depend on the number of GPUs. */
int mcb_price_european_packed(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                              int option_type, mcb_result *out);

/* Bullet (barrier-count) option, N_STEPS - Tk steps, restartable from (Ik, Sk, Tk).
 * Replaces wrapper_gpu_bullet_option and wrapper_gpu_bullet_option_atomic
 * (inc/wrappers.cuh:59-93, 95-125) = simulateBulletOptionPriceMultipleBlockGPU[atomic]
 * (inc/trajectories.cuh:115-191, 193-271). */
int mcb_price_bullet(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                     int Ik, float Sk, int Tk, mcb_result *out);

/* Full trajectories, path-major: prices[(p-first_path)*N_STEPS + i] = S(t_{i+1}); counts
 * (nullable) the barrier count after that step.  Replaces
 * Simulation::simulate_outer_trajectories + its kernel (inc/testing.cuh:281-326, 46-73) and
 * simulate_outer_trajectories (inc/trajectories.cuh:273-351).  `where` says whether prices /
 * counts are host or device pointers. */
int mcb_simulate_trajectories(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path,
                              uint64_t n_paths, uint64_t seed, float *prices, int *counts, int where);

/* Nested Monte Carlo: F[(p-first_outer)*N_STEPS + k] = discount * mean over N_PATHS_INNER
 * inner paths started from the outer state (S, I) of path p at step k.  Replaces the three
 * wrappers wrapper_gpu_bullet_option_nmc_{one_point_one_block,one_kernel,optimal}
 * (inc/wrappers.cuh:128-206, 209-266, 268-340) and their kernels (inc/nmc.cuh:12-386).
 * prices / counts (nullable) receive the outer trajectories.  *mean_F (nullable) receives
 * the wrappers' diagnostic scalar: sum(F) / (n_outer*N_STEPS + 1) (inc/wrappers.cuh:185-189). */
int mcb_nested_monte_carlo(mcb_engine *e, const mcb_option_data *opt, uint64_t first_outer,
                           uint64_t n_outer, uint64_t seed_outer, uint64_t seed_inner,
                           int discount_mode, float *F, float *prices, int *counts, int where,
                           double *mean_F);

/* Batched strike/vol sweep with common random numbers (BASELINE config 5): out[i] is
 * bit-identical to mcb_price_european with K = strikes[i], v = vols[i].  New capability. */
int mcb_price_sweep(mcb_engine *e, const mcb_option_data *opt, const float *strikes, const float *vols,
                    int n_params, uint64_t n_paths, uint64_t seed, int option_type, mcb_result *out);

/* Deterministic float sum of an array (replaces reduce3..6, inc/reduce.cuh:9-227, as used
 * by Simulation::test_reduction, inc/testing.cuh:185-235). */
int mcb_reduce_sum(mcb_engine *e, const float *x, uint64_t n, int where, float *out);

/* Per-block float sums with the reference's reduce3..6 index ranges (inc/reduce.cuh:9-227) as
 * Simulation::test_reduction drives them (inc/testing.cuh:185-235): block b of n_blocks sums
 * x[(b + k*n_blocks)*span .. +span) for k = 0 only (strided == 0: reduce3/4/5, span = 2*threads)
 * or for every k (strided != 0: reduce6's grid-stride loop).  Each block sum goes through the
 * engine's fixed 256-slot tree (slot t adds its elements of the block's index list in order),
 * so it is deterministic; out (host array) receives n_blocks floats. */
int mcb_reduce_blocks(mcb_engine *e, const float *x, uint64_t n, int where, uint32_t n_blocks, uint64_t span,
                      int strided, float *out);

/* n standard normals: out[i] = normal i of the stream (seed, subsequence 0), i.e. what
 * curand_normal would return from curand_init(seed, 0, 0, Philox).  Replaces
 * generate_random_array / init_random_array (cuRAND host API, inc/testing.cuh:17-42). */
int mcb_generate_normals(mcb_engine *e, uint64_t seed, uint64_t n, float *out, int where);

/* The reference's only on-disk format (testing.cu:37-47): header "time,trajectory,value", then per
 * trajectory a t = 0 row with x0 followed by ((1+i)*dt, trajectory, prices[p*n_steps + i]).
 * prices is a HOST array (e.g. from mcb_simulate_trajectories).  Pure host code. */
int mcb_write_trajectories_csv(const char *path, const float *prices, uint64_t n_trajectories, int n_steps,
                               float x0, float dt);

/* European-style pricing from pre-generated normals normals[p*n_steps + i]; payoffs[p].
 * Replaces simulateOptionPriceGPU / simulateOptionPriceMultipleBlockGPU
 * (inc/trajectories.cuh:14-52) and the CPU overload (inc/testing.cuh:75-91). */
int mcb_price_from_normals(mcb_engine *e, const mcb_option_data *opt, const float *normals,
                           uint64_t n_paths, int n_steps, float *payoffs, int where);

/* ---- sharded, stream-ordered pieces (one process per GPU; no host sync inside) -------
 * rank g of `world` owns segments [g*64/world, (g+1)*64/world).  The *_segments_async
 * calls fill d_segments[MCB_SEGMENTS][2] (sum, sumsq; device doubles) with the owned
 * segments and +0.0 elsewhere, so ONE sum-allreduce (or all-gather) of 1 KiB makes every
 * rank hold all 64; mcb_combine_segments_async then runs the fixed final tree.  With
 * world == 1 no collective is needed.  `stream` is a cudaStream_t (NULL = engine stream). */
int mcb_european_segments_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths,
                                uint64_t seed, int option_type, int rank, int world,
                                double *d_segments, void *stream);
int mcb_bullet_segments_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths,
                              uint64_t seed, int Ik, float Sk, int Tk, int rank, int world,
                              double *d_segments, void *stream);
/* strikes / vols are HOST arrays (n_params floats each); d_segments holds n_params sets of
 * [MCB_SEGMENTS][2] device doubles. */
int mcb_sweep_segments_async(mcb_engine *e, const mcb_option_data *opt, const float *strikes,
                             const float *vols, int n_params, uint64_t n_paths, uint64_t seed,
                             int option_type, int rank, int world, double *d_segments, void *stream);
int mcb_combine_segments_async(mcb_engine *e, const double *d_segments, int n_sets, uint64_t n_paths,
                               float r, float T, mcb_result *d_results, void *stream);
/* Same as mcb_simulate_trajectories with device pointers, but only enqueues. */
int mcb_trajectories_async(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path,
                           uint64_t n_paths, uint64_t seed, float *d_prices, int *d_counts, void *stream);
int mcb_nested_async(mcb_engine *e, const mcb_option_data *opt, uint64_t first_outer, uint64_t n_outer,
                     uint64_t seed_outer, uint64_t seed_inner, int discount_mode, float *d_F,
                     float *d_prices, int *d_counts, void *stream);

/* ---- the European job pipeline: ONE launch per price, results through mapped host memory --------
 * (a shard of 2^26 paths or more prices with the plain kernel plus one small segment launch: the ticket that
 * makes the single launch possible costs more than a launch there)
 * mcb_price_european is submit + collect.  A job is priced by the engine's group of shards:
 *   - a plain engine: one shard.  The pricing kernel's last CTA folds the segments, runs the final
 *     tree and writes the mcb_result into pinned host memory: one launch, no copy, no stream sync
 *     (replaces setup_kernel + pricing kernel + cudaDeviceSynchronize + cudaMemcpy + host finalise of
 *     inc/wrappers.cuh:38-56);
 *   - a multi-device engine (mcb_engine_create_multi): one shard per listed GPU, one launch each;
 *   - one engine per PROCESS (one process per GPU, e.g. under torchrun), connected with
 *     mcb_peer_mailbox_create / _connect below: rank g's kernel stores the segments it owns into
 *     every rank's mailbox over NVLink (CUDA-IPC mapped peer memory) and publishes an epoch flag.
 * With more than one shard the final tree is a one-warp kernel on the engine's second stream that
 * waits for the flags (bounded, see mcb_set_wait_timeout_ms), so the pricing stream never waits for a
 * peer: job e + 1 is priced while job e's segments are still crossing NVLink.  Up to
 * MCB_PIPELINE_DEPTH jobs may be in flight; the last MCB_RESULT_RING results can be collected.
 * A job of at most 2^20 paths (the reference's own call sizes, hello.cu:14) is latency-bound and is NOT
 * sharded: the leader of a multi-device engine / every rank of a process group prices it alone, with a
 * cluster of eight CTAs per 16384-path chunk (10-14 us per synchronous call; the tree is the same, so are
 * the bits).
 * Results are bit-identical for every group shape.  Every rank of a process group must submit the
 * same jobs in the same order.  If a peer never delivers, collect returns MCB_ERR_TIMEOUT and the
 * result is NaN with n_paths = 0 (never a plausible number). */
#define MCB_PIPELINE_DEPTH 4
#define MCB_RESULT_RING 8
int mcb_european_submit(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                        int option_type, uint64_t *ticket);
int mcb_european_collect(mcb_engine *e, uint64_t ticket, mcb_result *out);
/* Device time of a batch of pipelined jobs: start drains every stream and records an event,
 * stop records one behind everything submitted since (all shards, all of their streams), waits for it and
 * returns the elapsed milliseconds (CUDA events on the launching streams). */
int mcb_pipeline_timer_start(mcb_engine *e);
int mcb_pipeline_timer_stop(mcb_engine *e, double *elapsed_ms);

/* One engine per process: this rank's mailbox as a 64-byte cudaIpcMemHandle; the caller all-gathers
 * the handles and the ranks' mcb_peer_epoch values (e.g. torch.distributed.all_gather_object) and
 * passes all handles plus the MAXIMUM epoch to connect, then runs a barrier before the first
 * submit.  Any world size up to MCB_MAX_PEERS; at most one rank per GPU (kernels of different
 * ranks wait for one another). */
#define MCB_IPC_HANDLE_BYTES 64
#define MCB_MAX_PEERS 16
int mcb_peer_mailbox_create(mcb_engine *e, void *handle_out);
int mcb_peer_epoch(mcb_engine *e, uint64_t *epoch);
int mcb_peer_mailbox_connect(mcb_engine *e, int rank, int world, const void *all_handles, uint64_t base_epoch);
/* Bound of every device-side wait for a peer (default 10 000 ms, measured with %globaltimer) and
 * the number of waits that ran out so far. */
int mcb_set_wait_timeout_ms(mcb_engine *e, uint64_t ms);
int mcb_peer_timeouts(mcb_engine *e, uint64_t *count);

/* Number of kernel launches this engine has issued (bench.py's gpu_launches). */
uint64_t mcb_launch_count(mcb_engine *e);

/* Per-kernel device timing for the roofline figures (no reference counterpart: the reference
 * has no timers at all).  While enabled, every launch of a hot-path
 * kernel is bracketed by a CUDA-event pair recorded on the launching stream;
 * mcb_timing_read waits for the recorded launches of `kernel`, returns their summed device
 * time and count, and forgets them. */
enum { MCB_KERNEL_EUROPEAN = 0, MCB_KERNEL_BULLET = 1, MCB_KERNEL_TRAJECTORY = 2, MCB_KERNEL_NESTED = 3,
       MCB_KERNEL_SWEEP = 4, MCB_KERNEL_EUROPEAN_PACKED = 5, MCB_KERNEL_COUNT = 6 };
int mcb_timing_enable(mcb_engine *e, int on);
int mcb_timing_read(mcb_engine *e, int kernel, double *total_ms, uint64_t *launches);

/* ---- parity hooks (used by tests/ to pin the integer stream and the reduction tree) -- */
/* words[4*i..] = Philox block `blocks[i]` of subsequence `subsequences[i]` (host arrays). */
int mcb_philox_blocks(mcb_engine *e, uint64_t seed, const uint64_t *subsequences, const uint64_t *blocks,
                      uint64_t n, uint32_t *words);
/* Same words, produced on the device by the cuRAND LIBRARY's curandStatePhilox4_32_10_t
 * (curand_init(seed, subsequence, 4*block) + curand4): the live oracle of SURVEY.md 8(c). */
int mcb_curand_blocks(mcb_engine *e, uint64_t seed, const uint64_t *subsequences, const uint64_t *blocks,
                      uint64_t n, uint32_t *words);
/* Exhaustive accuracy scan of the engine's MUFU Box-Muller pieces against double precision over
 * the 32-bit words [first_word, first_word + count): which = 0 radius sqrt(-2 ln u(x)), 1 sin of
 * the angle v(y), 2 cos.  Returns the largest absolute error and the number of non-finite (or,
 * for the radius, negative) results.  count = 2^32 covers every possible word. */
int mcb_boxmuller_scan(mcb_engine *e, int which, uint64_t first_word, uint64_t count, double *max_abs_error,
                       uint64_t *n_bad);
/* The engine's float normals n0..n0+count-1 of one stream (host array out). */
int mcb_stream_normals(mcb_engine *e, uint64_t seed, uint64_t subsequence, uint64_t n0, uint64_t count,
                       float *normals);
/* Per-path European payoffs (host array, n_paths floats) and per-chunk float partials
 * (host array, 2 floats per chunk) of paths [0, n_paths). */
int mcb_european_payoffs(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                         uint64_t seed, int option_type, float *payoffs);
int mcb_european_chunk_partials(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                                int option_type, float *partials, uint64_t n_chunks);
/* Per-path payoffs under packed keying (host array, n_paths floats). */
int mcb_european_packed_payoffs(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                                uint64_t seed, int option_type, float *payoffs);
int mcb_bullet_payoffs(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                       uint64_t seed, int Ik, float Sk, int Tk, float *payoffs);
/* Last segments [MCB_SEGMENTS][2] computed by a whole-job call (host array of 128 doubles). */
int mcb_last_segments(mcb_engine *e, double *segments);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
