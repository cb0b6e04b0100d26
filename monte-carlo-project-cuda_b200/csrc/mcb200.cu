// mcb200.cu -- engine + C-ABI (include/mcb200.h) of the B200-native Monte Carlo pricer.
//
// Host side of the drop-in boundary: what the reference does inside every wrapper_* of
// inc/wrappers.cuh (cudaMalloc states/outputs -> setup_kernel -> kernel -> sync -> D2H ->
// host finalise -> cudaFree) becomes a persistent engine handle with its own stream and
// grow-only workspaces; parameters travel as kernel arguments (no __constant__ symbol the
// caller must remember to upload, cf. hello.cu:22); nothing prints, nothing exits.
// There is deliberately no CPU fallback anywhere in this file.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <curand_kernel.h>  // device API, used ONLY by the mcb_curand_blocks parity hook

#include "../../include/mcb200.h"
#include "path_kernels.cuh"
#include "pricing_kernels.cuh"

using namespace mcb;

static_assert(sizeof(mcb_option_data) == 48, "must match the reference's OptionData (inc/tool.cuh:13-26)");
static_assert(sizeof(mcb_result) == sizeof(ResultDev), "mcb_result layout");
static_assert(MCB_SLOTS == kSlots && MCB_SEGMENTS == kSegments, "reduction geometry");

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess)                                                                  \
            return fail(MCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),   \
                        __FILE__, __LINE__);                                                       \
    } while (0)

constexpr double kLog2e = 1.4426950408889634074;
constexpr double kSqrt2Ln2d = 1.1774100225154746910;  // sqrt(2 ln 2), see philox.cuh bm_radius_unscaled

template <typename T>
struct DeviceBuffer {
    T *ptr = nullptr;
    size_t cap = 0;  // elements
    int reserve(size_t n)
    {
        if (n <= cap) return MCB_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t err = cudaMalloc(&ptr, want * sizeof(T));
        if (err != cudaSuccess) {
            cudaGetLastError();
            return fail(MCB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(err));
        }
        cap = want;
        return MCB_OK;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}  // namespace

constexpr int kHostRing = 8;   // MCB_RESULT_RING: results kept for mcb_european_collect

// One launcher thread per non-leading shard of a multi-device engine: a European job is then enqueued
// on every GPU at the same time instead of one device after the other (8 devices x ~7 us of launch
// calls would start the last GPU ~50 us after the first -- 15 % of a 0.3 ms job).  A worker spins
// for a little while after each job (back-to-back submits find it hot) and then sleeps.
struct ShardWorker {
    std::thread thread;
    std::mutex mutex;
    std::condition_variable wake;
    std::atomic<unsigned long long> posted{0}, done{0};
    std::function<int()> job;
    int rc = 0;
    int device = 0;
    std::string message;
    bool quit = false;

    void loop()
    {
        cudaSetDevice(device);   // this thread only ever launches on its shard's device
        unsigned long long seen = 0;
        for (;;) {
            int spins = 0;
            while (posted.load(std::memory_order_acquire) == seen) {
                if (++spins < 200000) {
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                    continue;
                }
                std::unique_lock<std::mutex> lock(mutex);
                wake.wait(lock, [&] { return quit || posted.load(std::memory_order_acquire) != seen; });
                if (quit) return;
            }
            if (quit) return;
            seen = posted.load(std::memory_order_acquire);
            rc = job();
            done.store(seen, std::memory_order_release);
        }
    }
    void post(std::function<int()> fn)
    {
        job = std::move(fn);
        {
            std::lock_guard<std::mutex> lock(mutex);   // pairs with the predicate check of a sleeping worker
            posted.fetch_add(1, std::memory_order_release);
        }
        wake.notify_one();
    }
    int wait()
    {
        const unsigned long long want = posted.load(std::memory_order_acquire);
        while (done.load(std::memory_order_acquire) != want) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        return rc;
    }
    void stop()
    {
        {
            std::lock_guard<std::mutex> lock(mutex);
            quit = true;
            posted.fetch_add(1, std::memory_order_release);
        }
        wake.notify_one();
        if (thread.joinable()) thread.join();
    }
};

struct mcb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;         // main stream: every pricing / trajectory launch
    cudaStream_t stream2 = nullptr;        // odd European jobs of a world > 1 group: job e + 1 fills the SMs that
                                           // job e's tail, segment launch and events leave idle
    cudaStream_t f_stream = nullptr;       // final passes of world > 1 jobs (never blocks the pricing streams)
    cudaDeviceProp prop{};
    DeviceBuffer<float2> partials;
    DeviceBuffer<float2> partials2;        // chunk partials of the jobs on stream2
    DeviceBuffer<double> segments;
    DeviceBuffer<ResultDev> results;
    DeviceBuffer<unsigned char> scratch;   // hooks / host<->device staging
    DeviceBuffer<float> nested_ws;         // nested MC: log2 S, (prices), (counts) of the outer points
    DeviceBuffer<float> traj_ws;           // mcb_simulate_trajectories to a host buffer: one slab of rows (+ counts)
    DeviceBuffer<float4> sweep_sets;       // sweep: (c0, c1, K, -) per parameter set
    std::vector<float4> h_sweep_sets;      // host staging for sweep_sets (pageable on purpose)
    // ---- the job pipeline (mcb_european_submit / collect) ----
    PeerMailbox *mailbox = nullptr;        // this shard's mailbox (its own HBM)
    PeerTable peers{};                     // box[r] = shard r's mailbox as addressable from this device
    unsigned int *seg_tickets = nullptr;   // 2 x [kSegments + 1] (one set per pricing stream), zero between launches
    unsigned long long last_job_epoch = 0; // mcb_last_segments: the last collected European job ...
    uint64_t last_job_chunks = 0;          // ... and its chunk count (0: the last whole job left h_segments instead)
    uint64_t ring_chunks[kHostRing] = {};  // chunk count of the job in each host slot
    bool ring_own[kHostRing] = {};         // ... and whether it was priced as a group of one (segments in mailbox->own)
    bool last_job_own = false;
    unsigned long long slot_last_sharded[kRing] = {};   // separate processes: last sharded job per mailbox slot (acks)
    int rank = 0, world = 1;               // this shard's place in its group
    bool in_process = false;               // group = the shards of ONE multi-device engine (events, no device spins)
    bool ipc = false;                      // group = one engine per process, mailboxes mapped over CUDA IPC
    void *peer_mapped[kMaxPeers] = {};     // what cudaIpcOpenMemHandle returned (to close on destroy)
    std::vector<mcb_engine *> shards;      // leader of a multi-device engine: every shard, itself first
    ShardWorker *worker = nullptr;         // non-leading shard of a multi-device engine: its launcher thread
    mcb_engine *leader = nullptr;          // sub-engine of a multi-device engine: its leader
    unsigned long long job_epoch = 0;      // jobs submitted so far (leader / single engine)
    unsigned long long timeout_ns = 10ull * 1000000000ull;   // bound of every device-side wait
    bool small_jobs = true;                // jobs of <= 64 chunks take the cluster kernel (MCB_SMALL_JOBS=0: tools only)
    cudaEvent_t p_done[kRing] = {};        // pricing launch of the job in each ring slot has finished (this shard)
    cudaEvent_t f_done[kRing] = {};        // final pass of the job in each ring slot has finished (leader)
    cudaEvent_t t_begin = nullptr, t_end = nullptr;   // mcb_pipeline_timer_*
    HostSlot *h_ring = nullptr;            // mapped pinned: results of the last kHostRing jobs
    mcb_result *h_results = nullptr;       // pinned
    size_t h_results_cap = 0;
    double *h_segments = nullptr;          // mapped pinned, [MCB_SEGMENTS][2] of the last whole-job call
    uint64_t launches = 0;
    // optional per-kernel CUDA-event timing (mcb_timing_enable): one (start, stop) pair per
    // hot-path launch, recorded on the launching stream, read back by mcb_timing_read.
    bool timing = false;
    struct TimedLaunch {
        cudaEvent_t start, stop;
        int kernel;
    };
    std::vector<TimedLaunch> timed;      // recorded since the last read
    std::vector<TimedLaunch> event_pool; // recycled pairs
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

cudaStream_t pick(mcb_engine *e, void *stream) { return stream ? (cudaStream_t)stream : e->stream; }

// Brackets one kernel launch with CUDA events on the launching stream when timing is on.
constexpr size_t kMaxTimedLaunches = 1u << 16;
struct TimedScope {
    mcb_engine *e;
    cudaStream_t st;
    cudaEvent_t stop = nullptr;
    TimedScope(mcb_engine *eng, int kernel, cudaStream_t stream) : e(eng), st(stream)
    {
        if (!e->timing || e->timed.size() >= kMaxTimedLaunches) return;
        mcb_engine::TimedLaunch t{};
        if (!e->event_pool.empty()) {
            t = e->event_pool.back();
            e->event_pool.pop_back();
        } else if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        t.kernel = kernel;
        cudaEventRecord(t.start, st);
        stop = t.stop;
        e->timed.push_back(t);
    }
    ~TimedScope()
    {
        if (stop) cudaEventRecord(stop, st);
    }
};

int check_common(const mcb_engine *e, const mcb_option_data *o)
{
    if (!e) return fail(MCB_ERR_INVALID, "engine is NULL");
    if (!o) return fail(MCB_ERR_INVALID, "option data is NULL");
    if (!(o->S0 > 0.0f) || !std::isfinite(o->S0)) return fail(MCB_ERR_INVALID, "S0 must be positive and finite");
    if (!std::isfinite(o->K)) return fail(MCB_ERR_INVALID, "K must be finite");
    if (!(o->T > 0.0f) || !std::isfinite(o->T)) return fail(MCB_ERR_INVALID, "T must be positive and finite");
    if (!(o->v >= 0.0f) || !std::isfinite(o->v)) return fail(MCB_ERR_INVALID, "v (sigma) must be >= 0 and finite");
    if (!std::isfinite(o->r)) return fail(MCB_ERR_INVALID, "r must be finite");
    return MCB_OK;
}

int check_walk(const mcb_option_data *o)
{
    if (o->N_STEPS < 1) return fail(MCB_ERR_INVALID, "N_STEPS must be >= 1");
    if (!(o->step > 0.0f) || !std::isfinite(o->step)) return fail(MCB_ERR_INVALID, "step (dt) must be positive");
    if (!(o->v > 0.0f)) return fail(MCB_ERR_INVALID, "v (sigma) must be > 0 for multi-step walks");
    return MCB_OK;
}

// St = 2^(c0 + c1 z): constants folded in double, rounded once to float.
EuropeanParams european_params(const mcb_option_data *o, float K, float sigma, uint64_t n_paths_end, uint64_t seed,
                               uint64_t first_chunk)
{
    EuropeanParams p{};
    const double S0 = o->S0, r = o->r, sig = sigma, T = o->T;
    p.c0 = (float)(std::log2(S0) + (r - 0.5 * sig * sig) * T * kLog2e);
    p.c1 = (float)(sig * std::sqrt(T) * kLog2e * kSqrt2Ln2d);  // times the UNSCALED Box-Muller radius
    p.K = K;
    p.n_paths = n_paths_end;
    p.first_chunk = first_chunk;
    p.keys = make_philox_keys(seed);
    return p;
}

struct WalkConsts {
    float l0, sc, dr, v, lB;
    double l0d, scd, drd, lBd;   // the same in double (lBd = -inf without a barrier)
};

WalkConsts walk_consts(const mcb_option_data *o, double start)
{
    WalkConsts w;
    const double r = o->r, sig = o->v, dt = o->step;
    w.l0d = std::log2(start);
    w.scd = sig * std::sqrt(dt) * kLog2e * kSqrt2Ln2d;
    w.drd = (r - 0.5 * sig * sig) * dt * kLog2e;
    w.lBd = o->B > 0.0f ? std::log2((double)o->B) : -INFINITY;
    w.l0 = (float)w.l0d;
    w.sc = (float)w.scd;
    w.dr = (float)w.drd;
    w.v = (float)(sig * std::sqrt(dt) * kLog2e);
    w.lB = (float)w.lBd;
    return w;
}

// Threshold table of the walk kernels in dynamic shared memory: n_steps floats rounded up to 4, when it fits.
constexpr int kWalkTableMaxSteps = 8192;   // 32 KiB

void segment_span(int rank, int world, uint64_t n_chunks, int *seg_lo, int *seg_hi, uint64_t *chunk_lo,
                  uint64_t *chunk_hi)
{
    *seg_lo = (int)(((int64_t)rank * MCB_SEGMENTS) / world);
    *seg_hi = (int)(((int64_t)(rank + 1) * MCB_SEGMENTS) / world);
    *chunk_lo = (n_chunks * (uint64_t)*seg_lo) / MCB_SEGMENTS;
    *chunk_hi = (n_chunks * (uint64_t)*seg_hi) / MCB_SEGMENTS;
}

int check_shard(int rank, int world)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(MCB_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    return MCB_OK;
}

template <int PPS>
int launch_european(mcb_engine *e, const EuropeanParams &prm, int option_type, uint64_t n_ctas, float2 *partials,
                    float *payoffs, uint64_t payoffs_first, cudaStream_t st)
{
    if (n_ctas == 0) return MCB_OK;
    if (n_ctas > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    {
        TimedScope timed(e, MCB_KERNEL_EUROPEAN, st);
        if (option_type == MCB_PUT)
            european_kernel<kPut, PPS><<<(unsigned)n_ctas, kSlots, 0, st>>>(prm, partials, payoffs, payoffs_first);
        else
            european_kernel<kCall, PPS><<<(unsigned)n_ctas, kSlots, 0, st>>>(prm, partials, payoffs, payoffs_first);
    }
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

int launch_segments(mcb_engine *e, const float2 *partials, uint64_t stride, uint64_t first_chunk, uint64_t n_chunks,
                    int seg_lo, int seg_hi, int n_sets, double *d_segments, cudaStream_t st, int write_unowned = 1)
{
    if (n_sets >= 8) {   // many small folds (the sweep): one warp per (set, segment), same tree
        const uint64_t items = (uint64_t)n_sets * MCB_SEGMENTS;
        segment_sets_kernel<<<(unsigned)((items + kWarps - 1) / kWarps), kSlots, 0, st>>>(
            partials, stride, first_chunk, n_chunks, seg_lo, seg_hi, write_unowned, n_sets, d_segments);
    } else {
        segment_kernel<<<dim3(MCB_SEGMENTS, (unsigned)n_sets), kSlots, 0, st>>>(partials, stride, first_chunk, n_chunks,
                                                                               seg_lo, seg_hi, write_unowned, d_segments);
    }
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

int reserve_results(mcb_engine *e, size_t n)
{
    int rc = e->results.reserve(n);
    if (rc) return rc;
    if (n > e->h_results_cap) {
        if (e->h_results) cudaFreeHost(e->h_results);
        e->h_results = nullptr;
        e->h_results_cap = 0;
        CU(cudaMallocHost(&e->h_results, (n + 16) * sizeof(mcb_result)));
        e->h_results_cap = n + 16;
    }
    return MCB_OK;
}

constexpr uint64_t kEuropeanChunk = (uint64_t)MCB_SLOTS * MCB_EUROPEAN_PATHS_PER_SLOT;
constexpr uint64_t kBulletChunk = (uint64_t)MCB_SLOTS * MCB_BULLET_PATHS_PER_SLOT;
// European jobs whose shard has at least this many chunks (2^26 paths, ~0.16 ms) price with the plain
// kernel + a segment launch instead of the single fused launch (see segments_job_kernel)
constexpr uint64_t kTwoLaunchChunks = 4096;

uint64_t resolve_paths(const mcb_option_data *o, uint64_t n_paths)
{
    return n_paths ? n_paths : (o->N_PATHS > 0 ? (uint64_t)o->N_PATHS : 0);
}

__global__ void curand_blocks_kernel(uint64_t seed, const uint64_t *__restrict__ subseq,
                                     const uint64_t *__restrict__ block, uint64_t n, uint4 *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    curandStatePhilox4_32_10_t s;
    curand_init(seed, subseq[i], 4ull * block[i], &s);
    out[i] = curand4(&s);
}

// Tuning knobs of the trajectory launcher (tools/traj_bench.py sweeps them through the environment;
// unset = the shipped configuration).
int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// One row layout (SPL steps per lane, LPR lanes per row): the TMA slab kernel when the row fits one
// pass and is 16-byte aligned (the bandwidth path, config 3; counts and log2 prices ride along in
// their own staging rows), the general kernel otherwise.
template <int SPL, int LPR>
void launch_trajectory(const PathParams &prm_in, uint64_t n_paths, bool vec, bool base_aligned, float *d_prices,
                       int *d_counts, float *d_logs, cudaStream_t st)
{
    constexpr int kRowsPerWarp = 32 / LPR;
    constexpr int kSlabWarps = 4;
    const PathParams &prm = prm_in;
    // rows longer than one pass: one row group per slab with the whole rows staged while that leaves enough CTAs
    // per SM (<= 40 KB per CTA), pass-by-pass staging beyond (trajectory_long_kernel).  Measured on 2^18-2^20
    // rows (tools/traj_long_bench.py, profiles/r2_trajectory_tuning.txt), whole-row / pass-wise / general kernel:
    // prices 512 steps 4.15 / 3.62, 1024: 4.04 / 3.75, 1536: 3.63 / 3.72 / 3.63, 2048: - / 3.75 / 3.38, 4096: - / 3.78 /
    // 3.37 TB/s; prices + counts 512: 5.92 / 5.31, 1024: 4.58 / 5.23, 1536: - / 5.31 / 2.76, 4096: - / 5.29 / 2.75 TB/s.
    // Unaligned rows longer than a pass take the general kernel.
    const int n_arrays = 1 + (d_counts ? 1 : 0) + (d_logs ? 1 : 0);
    const size_t multi_smem = (size_t)kSlabWarps * n_arrays * kRowsPerWarp * (size_t)prm.n_steps * sizeof(float);
#define MCB_SLAB_W(ROWS, CNT, LOG, MULTI, ALIGNED, FAST, WARPS)                                               \
    do {                                                                                                      \
        auto kern = trajectory_slab_kernel<SPL, LPR, ROWS, WARPS, CNT, LOG, MULTI, ALIGNED, FAST>;            \
        const uint64_t rows_per_cta = (uint64_t)(WARPS) * (ROWS);                                             \
        const uint64_t ctas = (n_paths + rows_per_cta - 1) / rows_per_cta;                                    \
        const size_t smem = (size_t)(WARPS) * (1 + (CNT ? 1 : 0) + (LOG ? 1 : 0)) * (ROWS) *                  \
                            (size_t)prm.n_steps * sizeof(float);                                              \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<(unsigned)ctas, (WARPS) * 32, smem, st>>>(prm, d_prices, d_counts, d_logs);                    \
    } while (0)
#define MCB_SLAB(ROWS, CNT, LOG, MULTI, ALIGNED) MCB_SLAB_W(ROWS, CNT, LOG, MULTI, ALIGNED, false, kSlabWarps)
    // (only the largest layout is ever asked for rows longer than its pass)
    const int long_mode = env_int("MCB_TRAJ_LONG", 1);   // tools/: 0 general kernel, 2 pass-wise staging for every long row
    if (SPL * LPR == 256 && vec && prm.n_steps > SPL * LPR && multi_smem <= 40 * 1024 && long_mode != 2) {
        if constexpr (SPL * LPR == 256) {
            if (d_counts && d_logs) MCB_SLAB(kRowsPerWarp, true, true, true, true);
            else if (d_counts) MCB_SLAB(kRowsPerWarp, true, false, true, true);
            else if (d_logs) MCB_SLAB(kRowsPerWarp, false, true, true, true);
            else MCB_SLAB(kRowsPerWarp, false, false, true, true);
        }
    } else if (vec && prm.n_steps <= SPL * LPR) {
        // rows per slab: ~6 for one output array (tuned on B200 at 2^20 x 252, profiles/r1_trajectory_tuning.txt),
        // fewer when counts / logs need their own staging rows; always a whole number of passes
        constexpr int kRows1 = (6 + kRowsPerWarp - 1) / kRowsPerWarp * kRowsPerWarp;
        constexpr int kRows2 = (4 + kRowsPerWarp - 1) / kRowsPerWarp * kRowsPerWarp;
        constexpr int kRows3 = (2 + kRowsPerWarp - 1) / kRowsPerWarp * kRowsPerWarp;
        // hoisted Philox products + packed FP32x2 (the kernel's FAST flag): 2^20 x 252 prices 250 -> 238 us,
        // prices + counts 354 -> 344 us, same bits (profiles/r2_trajectory_tuning.txt); MCB_TRAJ_FAST=0 is
        // the plain form, kept for tools/traj_bench.py's A/B run
        const bool fast = env_int("MCB_TRAJ_FAST", 1) != 0;
        if (d_counts && d_logs) { if (fast) MCB_SLAB_W(kRows3, true, true, false, true, true, kSlabWarps); else MCB_SLAB(kRows3, true, true, false, true); }
        else if (d_counts) { if (fast) MCB_SLAB_W(kRows2, true, false, false, true, true, kSlabWarps); else MCB_SLAB(kRows2, true, false, false, true); }
        else if (d_logs) { if (fast) MCB_SLAB_W(kRows2, false, true, false, true, true, kSlabWarps); else MCB_SLAB(kRows2, false, true, false, true); }
        else { if (fast) MCB_SLAB_W(kRows1, false, false, false, true, true, kSlabWarps); else MCB_SLAB(kRows1, false, false, false, true); }
    } else if (base_aligned && prm.n_steps <= SPL * LPR) {
        // rows that are not a multiple of 4 floats (150, 250 steps ...): slabs of 8 / 4 rows start on
        // 16-byte boundaries, so the bulk store still applies; only the staging is element-wise
        if (d_counts && d_logs) MCB_SLAB(4, true, true, false, false);
        else if (d_counts) MCB_SLAB(4, true, false, false, false);
        else if (d_logs) MCB_SLAB(4, false, true, false, false);
        else MCB_SLAB(8, false, false, false, false);
#undef MCB_SLAB
#undef MCB_SLAB_W
    } else if (SPL * LPR == 256 && vec && prm.n_steps > SPL * LPR && long_mode != 0) {
        // rows too long to stage whole: pass-by-pass staging, two 1 KiB bulk stores per pass and array
        if constexpr (SPL * LPR == 256) {
            constexpr int kLongWarps = 4;
            const uint64_t pairs = (n_paths + 1) / 2;
            const unsigned ctas = (unsigned)((pairs + kLongWarps - 1) / kLongWarps);
            const size_t smem = (size_t)kLongWarps * 2 * n_arrays * 2 * 256 * sizeof(float);
#define MCB_LONG(CNT, LOG)                                                                                    \
            do {                                                                                              \
                auto kern = trajectory_long_kernel<CNT, LOG, kLongWarps>;                                     \
                if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                kern<<<ctas, kLongWarps * 32, smem, st>>>(prm, d_prices, d_counts, d_logs);                   \
            } while (0)
            if (d_counts && d_logs) MCB_LONG(true, true);
            else if (d_counts) MCB_LONG(true, false);
            else if (d_logs) MCB_LONG(false, true);
            else MCB_LONG(false, false);
#undef MCB_LONG
        }
    } else {
        const uint64_t rows_per_cta = (uint64_t)kPathWarps * kPathsPerWarp * kRowsPerWarp;
        const unsigned g = (unsigned)((n_paths + rows_per_cta - 1) / rows_per_cta), b = kPathWarps * 32;
        if (vec && d_counts) trajectory_kernel<SPL, LPR, kStoreVec4, true><<<g, b, 0, st>>>(prm, d_prices, d_counts, d_logs);
        else if (vec) trajectory_kernel<SPL, LPR, kStoreVec4, false><<<g, b, 0, st>>>(prm, d_prices, d_counts, d_logs);
        else if (d_counts) trajectory_kernel<SPL, LPR, kStoreScalar, true><<<g, b, 0, st>>>(prm, d_prices, d_counts, d_logs);
        else trajectory_kernel<SPL, LPR, kStoreScalar, false><<<g, b, 0, st>>>(prm, d_prices, d_counts, d_logs);
    }
}

}  // namespace

// ============================================================================================
extern "C" {

const char *mcb_last_error(void) { return g_error; }
int mcb_version(void) { return MCB_VERSION; }

static int engine_create_one(int device, mcb_engine **out)
{
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(MCB_ERR_NO_DEVICE, "no CUDA device: %s (this engine has no CPU fallback)",
                    err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    }
    if (device < 0 || device >= count) return fail(MCB_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    mcb_engine *e = new (std::nothrow) mcb_engine();
    if (!e) return fail(MCB_ERR_NOMEM, "out of host memory");
    e->device = device;
    DeviceGuard g(device);
    if (!g.ok) {
        delete e;
        return fail(MCB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    }
    err = cudaGetDeviceProperties(&e->prop, device);
    if (err == cudaSuccess && e->prop.major != 10) {
        int major = e->prop.major, minor = e->prop.minor;
        delete e;
        return fail(MCB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, major,
                    minor);
    }
    int prio_lo = 0, prio_hi = 0;
    if (err == cudaSuccess) err = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking);
    // the final passes are single warps that must not queue behind thousands of pricing CTAs
    if (err == cudaSuccess) err = cudaStreamCreateWithPriority(&e->f_stream, cudaStreamNonBlocking, prio_hi);
    if (err == cudaSuccess)
        err = cudaHostAlloc(&e->h_segments, sizeof(double) * (2 * MCB_SEGMENTS + 8),
                            cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess)
        err = cudaHostAlloc(&e->h_ring, sizeof(HostSlot) * kHostRing, cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) err = cudaMalloc(&e->mailbox, sizeof(PeerMailbox));
    if (err == cudaSuccess) err = cudaMemset(e->mailbox, 0, sizeof(PeerMailbox));
    if (err == cudaSuccess) err = cudaMalloc(&e->seg_tickets, sizeof(unsigned int) * 2 * (kSegments + 1));
    if (err == cudaSuccess) err = cudaMemset(e->seg_tickets, 0, sizeof(unsigned int) * 2 * (kSegments + 1));
    for (int i = 0; i < kRing && err == cudaSuccess; ++i) {
        err = cudaEventCreateWithFlags(&e->p_done[i], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->f_done[i], cudaEventDisableTiming);
    }
    if (err == cudaSuccess) err = cudaEventCreate(&e->t_begin);
    if (err == cudaSuccess) err = cudaEventCreate(&e->t_end);
    if (err != cudaSuccess) {
        int rc = fail(MCB_ERR_CUDA, "engine setup failed: %s", cudaGetErrorString(err));
        mcb_engine_destroy(e);
        return rc;
    }
    memset(e->h_segments, 0, sizeof(double) * (2 * MCB_SEGMENTS + 8));
    memset(e->h_ring, 0, sizeof(HostSlot) * kHostRing);
    e->small_jobs = env_int("MCB_SMALL_JOBS", 1) != 0;
    e->peers.box[0] = e->mailbox;
    if (e->segments.reserve(2 * MCB_SEGMENTS + 8) || reserve_results(e, 1)) {
        mcb_engine_destroy(e);
        return MCB_ERR_NOMEM;
    }
    *out = e;
    return MCB_OK;
}

int mcb_engine_create(int device, mcb_engine **out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    return engine_create_one(device, out);
}

int mcb_engine_create_multi(const int *devices, int n_devices, mcb_engine **out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > kMaxPeers)
        return fail(MCB_ERR_INVALID, "need 1..%d devices", kMaxPeers);
    std::vector<mcb_engine *> shards;
    int rc = MCB_OK;
    for (int i = 0; i < n_devices && rc == MCB_OK; ++i) {
        mcb_engine *s = nullptr;
        rc = engine_create_one(devices[i], &s);
        if (rc == MCB_OK) shards.push_back(s);
    }
    // every shard stores into the leader's mailbox / segment buffer: peer access both ways
    for (size_t i = 0; i < shards.size() && rc == MCB_OK; ++i)
        for (size_t j = 0; j < shards.size() && rc == MCB_OK; ++j) {
            const int a = shards[i]->device, b = shards[j]->device;
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (!can) {
                rc = fail(MCB_ERR_CUDA, "device %d cannot access device %d over NVLink/PCIe peer memory", a, b);
                break;
            }
            DeviceGuard g(a);
            cudaError_t err = cudaDeviceEnablePeerAccess(b, 0);
            if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (err != cudaSuccess) rc = fail(MCB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", a, b,
                                                   cudaGetErrorString(err));
        }
    if (rc != MCB_OK) {
        const std::string msg = g_error;   // the destructors below must not clobber the message
        for (mcb_engine *s : shards) mcb_engine_destroy(s);
        snprintf(g_error, sizeof(g_error), "%s", msg.c_str());
        return rc;
    }
    mcb_engine *L = shards[0];
    if (shards.size() > 1) {
        for (size_t i = 0; i < shards.size(); ++i) {
            mcb_engine *s = shards[i];
            s->rank = (int)i;
            s->world = (int)shards.size();
            s->in_process = true;
            s->leader = i ? L : nullptr;
            for (size_t r = 0; r < shards.size(); ++r) s->peers.box[r] = shards[r]->mailbox;
            if (i) {
                s->worker = new (std::nothrow) ShardWorker();
                if (s->worker) {
                    s->worker->device = s->device;
                    s->worker->thread = std::thread([w = s->worker] { w->loop(); });
                }
            }
        }
        L->shards = shards;
    }
    *out = L;
    return MCB_OK;
}

int mcb_engine_shard_count(mcb_engine *e) { return e ? (e->shards.empty() ? 1 : (int)e->shards.size()) : 0; }

int mcb_engine_destroy(mcb_engine *e)
{
    if (!e) return MCB_OK;
    if (!e->shards.empty()) {            // leader: the other shards go first (their stores target this one)
        std::vector<mcb_engine *> subs(e->shards.begin() + 1, e->shards.end());
        e->shards.clear();
        for (mcb_engine *s : subs) {
            s->leader = nullptr;
            mcb_engine_destroy(s);
        }
    }
    if (e->worker) {
        e->worker->stop();
        delete e->worker;
        e->worker = nullptr;
    }
    DeviceGuard g(e->device);
    if (e->stream) {
        cudaStreamSynchronize(e->stream);
        if (e->stream2) cudaStreamSynchronize(e->stream2);
        if (e->f_stream) cudaStreamSynchronize(e->f_stream);
        cudaStreamDestroy(e->stream);
    }
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->f_stream) cudaStreamDestroy(e->f_stream);
    for (auto *v : {&e->timed, &e->event_pool})
        for (auto &t : *v) {
            cudaEventDestroy(t.start);
            cudaEventDestroy(t.stop);
        }
    for (int i = 0; i < kRing; ++i) {
        if (e->p_done[i]) cudaEventDestroy(e->p_done[i]);
        if (e->f_done[i]) cudaEventDestroy(e->f_done[i]);
    }
    if (e->t_begin) cudaEventDestroy(e->t_begin);
    if (e->t_end) cudaEventDestroy(e->t_end);
    e->partials.release();
    e->partials2.release();
    e->segments.release();
    e->results.release();
    e->scratch.release();
    e->nested_ws.release();
    e->traj_ws.release();
    e->sweep_sets.release();
    for (int r = 0; r < kMaxPeers; ++r)
        if (e->peer_mapped[r]) cudaIpcCloseMemHandle(e->peer_mapped[r]);
    if (e->mailbox) cudaFree(e->mailbox);
    if (e->seg_tickets) cudaFree(e->seg_tickets);
    if (e->h_results) cudaFreeHost(e->h_results);
    if (e->h_segments) cudaFreeHost(e->h_segments);
    if (e->h_ring) cudaFreeHost(e->h_ring);
    cudaGetLastError();
    delete e;
    return MCB_OK;
}

int mcb_get_device_info(mcb_engine *e, mcb_device_info *out)
{
    if (!e || !out) return fail(MCB_ERR_INVALID, "NULL argument");
    memset(out, 0, sizeof(*out));
    strncpy(out->name, e->prop.name, sizeof(out->name) - 1);
    out->sm_count = e->prop.multiProcessorCount;
    out->cc_major = e->prop.major;
    out->cc_minor = e->prop.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
    out->clock_khz = khz;
    out->total_mem = e->prop.totalGlobalMem;
    return MCB_OK;
}

static int sync_all(mcb_engine *e)
{
    const size_t n = e->shards.empty() ? 1 : e->shards.size();
    for (size_t i = 0; i < n; ++i) {
        mcb_engine *s = e->shards.empty() ? e : e->shards[i];
        DeviceGuard g(s->device);
        CU(cudaStreamSynchronize(s->stream));
        CU(cudaStreamSynchronize(s->stream2));
        CU(cudaStreamSynchronize(s->f_stream));
    }
    return MCB_OK;
}

int mcb_synchronize(mcb_engine *e)
{
    if (!e) return fail(MCB_ERR_INVALID, "engine is NULL");
    return sync_all(e);
}

// --------------------------------------------------------------- peer-memory exchange (NVLink)
static_assert(MCB_MAX_PEERS == kMaxPeers && MCB_IPC_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "peer ABI");
static_assert(MCB_PIPELINE_DEPTH == kRing && MCB_RESULT_RING == kHostRing, "pipeline ABI");
static_assert(sizeof(HostSlot) == 128, "one result slot per 128-byte line");

int mcb_peer_mailbox_create(mcb_engine *e, void *handle_out)
{
    if (!e || !handle_out) return fail(MCB_ERR_INVALID, "NULL argument");
    if (e->in_process || e->leader) return fail(MCB_ERR_INVALID, "a multi-device engine is its own group");
    DeviceGuard g(e->device);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->mailbox));
    memcpy(handle_out, &h, sizeof(h));
    return MCB_OK;
}

int mcb_peer_epoch(mcb_engine *e, uint64_t *epoch)
{
    if (!e || !epoch) return fail(MCB_ERR_INVALID, "NULL argument");
    *epoch = e->job_epoch;
    return MCB_OK;
}

int mcb_peer_mailbox_connect(mcb_engine *e, int rank, int world, const void *all_handles, uint64_t base_epoch)
{
    if (!e || !all_handles) return fail(MCB_ERR_INVALID, "NULL argument");
    if (e->in_process || e->leader) return fail(MCB_ERR_INVALID, "a multi-device engine is its own group");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(MCB_ERR_INVALID, "bad rank/world %d/%d (at most %d peers)", rank, world, kMaxPeers);
    if (base_epoch < e->job_epoch)
        return fail(MCB_ERR_INVALID, "base_epoch %llu is behind this engine's %llu: pass the maximum over all ranks",
                    (unsigned long long)base_epoch, e->job_epoch);
    DeviceGuard g(e->device);
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaStreamSynchronize(e->stream2));
    CU(cudaStreamSynchronize(e->f_stream));
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int r = 0; r < kMaxPeers; ++r) {
        if (e->peer_mapped[r]) {
            cudaIpcCloseMemHandle(e->peer_mapped[r]);
            e->peer_mapped[r] = nullptr;
        }
        e->peers.box[r] = nullptr;
    }
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            e->peers.box[r] = e->mailbox;
            continue;
        }
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess));
        e->peer_mapped[r] = p;
        e->peers.box[r] = static_cast<PeerMailbox *>(p);
    }
    // Every rank starts the group's jobs at the SAME epoch (the caller all-gathers mcb_peer_epoch and
    // passes the maximum): flags and acks left by earlier groups are <= base_epoch, so they can
    // neither satisfy a wait of a new job nor hold a producer back.  My own mailbox is reset here;
    // the caller's barrier after connect orders this before any peer's first store.
    std::vector<unsigned long long> acks(kMaxPeers, (unsigned long long)base_epoch);
    CU(cudaMemcpy(e->mailbox->consumed, acks.data(), sizeof(unsigned long long) * kMaxPeers, cudaMemcpyHostToDevice));
    e->job_epoch = base_epoch;
    for (int k = 0; k < kRing; ++k) e->slot_last_sharded[k] = 0;   // (the acks were just reset to base_epoch)
    e->rank = rank;
    e->world = world;
    e->ipc = world > 1;
    return MCB_OK;
}

int mcb_set_wait_timeout_ms(mcb_engine *e, uint64_t ms)
{
    if (!e || ms == 0) return fail(MCB_ERR_INVALID, "bad argument");
    const size_t n = e->shards.empty() ? 1 : e->shards.size();
    for (size_t i = 0; i < n; ++i) (e->shards.empty() ? e : e->shards[i])->timeout_ns = ms * 1000000ull;
    return MCB_OK;
}

int mcb_peer_timeouts(mcb_engine *e, uint64_t *count)
{
    if (!e || !count) return fail(MCB_ERR_INVALID, "NULL argument");
    uint64_t total = 0;
    const size_t n = e->shards.empty() ? 1 : e->shards.size();
    for (size_t i = 0; i < n; ++i) {
        mcb_engine *s = e->shards.empty() ? e : e->shards[i];
        DeviceGuard g(s->device);
        unsigned int t = 0;
        CU(cudaMemcpy(&t, &s->mailbox->timeouts, sizeof(t), cudaMemcpyDeviceToHost));
        total += t;
    }
    *count = total;
    return MCB_OK;
}

// ---------------------------------------------------------------- the European job pipeline
namespace {

// Host side of HostSlot's flag-in-data format (pricing_kernels.cuh): a result is complete once all ten words carry
// the ticket's tag; a slot is expired by giving every word a tag that cannot match.
bool host_slot_read(const volatile HostSlot *hs, unsigned long long ticket, ResultDev *out)
{
    const unsigned long long tag = ticket & 0xffffffffull;
    unsigned long long f[5];
    for (int k = 0; k < 5; ++k) {
        const unsigned long long lo = hs->w[2 * k], hi = hs->w[2 * k + 1];   // aligned 8-byte loads: never torn
        if ((lo >> 32) != tag || (hi >> 32) != tag) return false;
        f[k] = (lo & 0xffffffffull) | (hi << 32);
    }
    memcpy(out, f, sizeof(f));
    return true;
}

void host_slot_expire(HostSlot *hs, unsigned long long epoch)
{
    const unsigned long long never = (~epoch & 0xffffffffull) << 32;
    for (int k = 0; k < 10; ++k) *(volatile unsigned long long *)&hs->w[k] = never;
}

// One shard's launch of job `epoch`: european_job_kernel over the chunks it owns (or the publish-only
// kernel when it owns none), on the shard's main stream.  solo: the leader of an in-process group prices the whole
// (small) job by itself, as a group of one -- the tree does not depend on the sharding, so the bits are the same.
int submit_shard(mcb_engine *s, mcb_engine *L, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                 int option_type, unsigned long long epoch, bool solo = false)
{
    DeviceGuard g(s->device);
    if (!g.ok) return fail(MCB_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    const int rank = solo ? 0 : s->rank, world = solo ? 1 : s->world;
    const int slot = (int)(epoch % kRing);
    // groups of several shards alternate their jobs between two pricing streams: the next job's CTAs
    // take the SM slots this job's last wave, segment launch and event records leave idle
    const int lane = (s->world > 1) ? (int)(epoch & 1ull) : 0;
    cudaStream_t st = lane ? s->stream2 : s->stream;
    DeviceBuffer<float2> &partials = lane ? s->partials2 : s->partials;
    const uint64_t n_chunks = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    if (c_hi - c_lo > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    int rc;
    if (partials.cap < (size_t)(c_hi - c_lo) + 1) {
        CU(cudaStreamSynchronize(st));   // growing the workspace frees the old one: nothing may still use it
        if ((rc = partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    }
    // in-process groups pace themselves with events: the mailbox slot of job epoch - kRing must have
    // been folded by the leader before this job's stores land in it (separate processes use acks)
    if (s->in_process && epoch > (unsigned long long)kRing) CU(cudaStreamWaitEvent(st, L->f_done[slot], 0));
    JobArgs args{};
    args.n_chunks = n_chunks;
    args.n_paths = n_paths;
    args.discount = std::exp(-(double)opt->r * (double)opt->T);
    args.seg_tickets = s->seg_tickets + lane * (kSegments + 1);
    args.peers = s->peers;
    args.epoch = epoch;
    args.timeout_ns = s->timeout_ns;
    args.seg_lo = seg_lo;
    args.seg_hi = seg_hi;
    args.live_segments = 0;
    for (int sg = seg_lo; sg < seg_hi; ++sg)
        if ((n_chunks * (uint64_t)sg) / MCB_SEGMENTS != (n_chunks * (uint64_t)(sg + 1)) / MCB_SEGMENTS)
            ++args.live_segments;
    args.rank = rank;
    args.world = world;
    args.n_consumers = (s->ipc && !solo) ? world : 1;
    if (s->ipc && !solo) {
        args.ack_epoch = s->slot_last_sharded[slot];
        s->slot_last_sharded[slot] = epoch;
    }
    if (solo) args.peers.box[0] = s->mailbox;
    if (world == 1) {
        args.d_out = s->results.ptr;
        args.h_out = &s->h_ring[epoch % kHostRing];
    }
    if (c_hi - c_lo >= kTwoLaunchChunks) {
        // large shard: the plain pricing kernel, then one CTA per owned segment (see segments_job_kernel)
        const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
        if ((rc = launch_european<MCB_EUROPEAN_PATHS_PER_SLOT>(s, prm, option_type, c_hi - c_lo, partials.ptr, nullptr, 0,
                                                               st)))
            return rc;
        segments_job_kernel<<<(unsigned)(seg_hi - seg_lo), kSlots, 0, st>>>(args, partials.ptr, c_lo);
    } else if (c_hi > c_lo && n_chunks <= kSmallJobChunks && s->small_jobs && world == 1) {
        // small job: a cluster of eight CTAs per chunk (latency, see european_small_job_kernel)
        const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
        TimedScope timed(s, MCB_KERNEL_EUROPEAN, st);
        const unsigned grid = (unsigned)(c_hi - c_lo) * kSmallCluster;
        if (option_type == MCB_PUT)
            european_small_job_kernel<kPut, MCB_EUROPEAN_PATHS_PER_SLOT><<<grid, kSlots, 0, st>>>(prm, args, partials.ptr);
        else
            european_small_job_kernel<kCall, MCB_EUROPEAN_PATHS_PER_SLOT><<<grid, kSlots, 0, st>>>(prm, args, partials.ptr);
    } else if (c_hi > c_lo) {
        const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
        TimedScope timed(s, MCB_KERNEL_EUROPEAN, st);
        if (option_type == MCB_PUT)
            european_job_kernel<kPut, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(
                prm, args, partials.ptr);
        else
            european_job_kernel<kCall, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(
                prm, args, partials.ptr);
    } else {
        job_publish_empty_kernel<<<1, 32, 0, st>>>(args);
    }
    s->launches++;
    CU(cudaGetLastError());
    if (world > 1) CU(cudaEventRecord(s->p_done[slot], st));
    return MCB_OK;
}

size_t shard_count(const mcb_engine *e) { return e->shards.empty() ? 1 : e->shards.size(); }
mcb_engine *shard_at(mcb_engine *e, size_t i) { return e->shards.empty() ? e : e->shards[i]; }

}  // namespace

int mcb_european_submit(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int option_type,
                        uint64_t *ticket)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (e->leader) return fail(MCB_ERR_INVALID, "submit to the multi-device engine, not to one of its shards");
    if (!ticket) return fail(MCB_ERR_INVALID, "ticket is NULL");
    if (option_type != MCB_CALL && option_type != MCB_PUT) return fail(MCB_ERR_INVALID, "bad option_type");
    n_paths = resolve_paths(opt, n_paths);
    if (n_paths == 0) return fail(MCB_ERR_INVALID, "n_paths must be > 0");
    // the epoch advances before anything that can fail, so that every rank of a group stays in step
    const unsigned long long epoch = ++e->job_epoch;
    *ticket = epoch;
    HostSlot *hs = &e->h_ring[epoch % kHostRing];
    host_slot_expire(hs, epoch);      // the job that used this slot kHostRing tickets ago expires here
    e->ring_chunks[epoch % kHostRing] = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    const size_t n = shard_count(e);
    // A small job is not worth sharding (a 10^5-path call is 10 us on one GPU, 30-40 us once launcher threads,
    // cross-device events or peer flags and a final pass are involved): the leader of an in-process group -- every
    // rank of a group of processes -- prices it alone as a group of one.  Same tree, same bits; it touches no
    // mailbox slot a peer writes, and the epoch advances on every rank as for any other job.
    const bool small = e->small_jobs && e->ring_chunks[epoch % kHostRing] <= kSmallJobChunks;
    const bool solo = small && ((n > 1 && e->in_process) || e->ipc);
    e->ring_own[epoch % kHostRing] = small && (solo || e->world == 1);
    if (solo) return submit_shard(e, e, opt, n_paths, seed, option_type, epoch, true);
    // the other shards' launcher threads enqueue their devices while this thread does the leader's
    for (size_t i = 1; i < n; ++i) {
        mcb_engine *s = shard_at(e, i);
        if (!s->worker) continue;
        s->worker->post([=]() {
            const int r = submit_shard(s, e, opt, n_paths, seed, option_type, epoch);
            if (r != MCB_OK) s->worker->message = g_error;
            return r;
        });
    }
    rc = submit_shard(e, e, opt, n_paths, seed, option_type, epoch);
    for (size_t i = 1; i < n; ++i) {
        mcb_engine *s = shard_at(e, i);
        const int r = s->worker ? s->worker->wait() : submit_shard(s, e, opt, n_paths, seed, option_type, epoch);
        if (r != MCB_OK && rc == MCB_OK) {
            rc = s->worker ? fail(r, "shard %zu (device %d): %s", i, s->device, s->worker->message.c_str()) : r;
        }
    }
    if (rc) return rc;
    if (e->world > 1) {
        DeviceGuard g(e->device);
        const int slot = (int)(epoch % kRing);
        for (size_t i = 0; i < n; ++i) CU(cudaStreamWaitEvent(e->f_stream, shard_at(e, i)->p_done[slot], 0));
        combine_job_kernel<<<1, 32, 0, e->f_stream>>>(e->peers, e->rank, e->world, e->ipc ? 1 : 0, epoch, e->timeout_ns,
                                                      (n_paths + kEuropeanChunk - 1) / kEuropeanChunk, n_paths,
                                                      std::exp(-(double)opt->r * (double)opt->T), e->results.ptr, hs);
        e->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->f_done[slot], e->f_stream));
    }
    return MCB_OK;
}

int mcb_european_collect(mcb_engine *e, uint64_t ticket, mcb_result *out)
{
    if (!e || !out) return fail(MCB_ERR_INVALID, "NULL argument");
    if (ticket == 0 || ticket > e->job_epoch) return fail(MCB_ERR_INVALID, "unknown ticket %llu", (unsigned long long)ticket);
    if (e->job_epoch - ticket >= (unsigned long long)kHostRing)
        return fail(MCB_ERR_INVALID, "ticket %llu has expired (only the last %d results are kept)",
                    (unsigned long long)ticket, kHostRing);
    volatile HostSlot *hs = &e->h_ring[ticket % kHostRing];
    // The result arrives in mapped host memory straight from the kernel: spin until its ten tagged words are there
    // (no cudaStreamSynchronize round trip); every so often make sure the streams are still healthy.
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long long spins = 0;
    ResultDev got;
    while (!host_slot_read(hs, ticket, &got)) {
        if ((++spins & 0x3fff) == 0) {
            DeviceGuard g(e->device);
            cudaError_t qa = cudaStreamQuery(e->stream), qb = cudaStreamQuery(e->f_stream);
            const cudaError_t qc = cudaStreamQuery(e->stream2);
            if (qa == cudaSuccess) qa = qc;
            if ((qa != cudaSuccess && qa != cudaErrorNotReady) || (qb != cudaSuccess && qb != cudaErrorNotReady))
                return fail(MCB_ERR_CUDA, "stream failed while waiting for ticket %llu: %s", (unsigned long long)ticket,
                            cudaGetErrorString(qa != cudaSuccess && qa != cudaErrorNotReady ? qa : qb));
            if (qa == cudaSuccess && qb == cudaSuccess) {
                bool idle = true;       // multi-device: every shard's stream must have drained too
                for (size_t i = 1; i < shard_count(e) && idle; ++i) {
                    DeviceGuard gs(shard_at(e, i)->device);
                    idle = cudaStreamQuery(shard_at(e, i)->stream) == cudaSuccess &&
                           cudaStreamQuery(shard_at(e, i)->stream2) == cudaSuccess;
                }
                if (idle && !host_slot_read(hs, ticket, &got)) {
                    std::atomic_thread_fence(std::memory_order_seq_cst);
                    if (!host_slot_read(hs, ticket, &got))
                        return fail(MCB_ERR_CUDA, "the streams are idle but ticket %llu never completed",
                                    (unsigned long long)ticket);
                }
            }
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (waited > 600.0) return fail(MCB_ERR_TIMEOUT, "ticket %llu: no result after 600 s", (unsigned long long)ticket);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    memcpy(out, &got, sizeof(mcb_result));
    e->last_job_epoch = ticket;
    e->last_job_chunks = e->ring_chunks[ticket % kHostRing];
    e->last_job_own = e->ring_own[ticket % kHostRing];
    if (out->n_paths == 0 || out->price != out->price)
        return fail(MCB_ERR_TIMEOUT, "ticket %llu: a peer did not deliver its segments within %.1f s (result poisoned)",
                    (unsigned long long)ticket, (double)e->timeout_ns * 1e-9);
    return MCB_OK;
}

int mcb_pipeline_timer_start(mcb_engine *e)
{
    if (!e) return fail(MCB_ERR_INVALID, "engine is NULL");
    int rc = sync_all(e);
    if (rc) return rc;
    DeviceGuard g(e->device);
    CU(cudaEventRecord(e->t_begin, e->stream));
    // nothing of the timed jobs may start before t_begin, on any stream of any shard
    CU(cudaStreamWaitEvent(e->f_stream, e->t_begin, 0));
    for (size_t i = 0; i < shard_count(e); ++i) {
        DeviceGuard gs(shard_at(e, i)->device);
        if (i) CU(cudaStreamWaitEvent(shard_at(e, i)->stream, e->t_begin, 0));
        CU(cudaStreamWaitEvent(shard_at(e, i)->stream2, e->t_begin, 0));
    }
    return MCB_OK;
}

int mcb_pipeline_timer_stop(mcb_engine *e, double *elapsed_ms)
{
    if (!e || !elapsed_ms) return fail(MCB_ERR_INVALID, "NULL argument");
    {
        // the end of the timed region is when EVERY stream of every shard has drained
        DeviceGuard g(e->device);
        for (size_t i = 0; i < shard_count(e); ++i) {
            mcb_engine *s = shard_at(e, i);
            DeviceGuard gs(s->device);
            for (cudaStream_t st : {s->stream, s->stream2}) {
                cudaEvent_t tmp = nullptr;
                CU(cudaEventCreateWithFlags(&tmp, cudaEventDisableTiming));
                CU(cudaEventRecord(tmp, st));
                CU(cudaStreamWaitEvent(e->f_stream, tmp, 0));
                CU(cudaEventDestroy(tmp));   // released once the wait has consumed it
            }
        }
        CU(cudaEventRecord(e->t_end, e->f_stream));
        CU(cudaEventSynchronize(e->t_end));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, e->t_begin, e->t_end));
        *elapsed_ms = (double)ms;
    }
    return MCB_OK;
}

uint64_t mcb_launch_count(mcb_engine *e)
{
    if (!e) return 0;
    uint64_t total = 0;
    for (size_t i = 0; i < shard_count(e); ++i) total += shard_at(e, i)->launches;
    return total;
}

int mcb_timing_enable(mcb_engine *e, int on)
{
    if (!e) return fail(MCB_ERR_INVALID, "engine is NULL");
    e->timing = on != 0;
    return MCB_OK;
}

int mcb_timing_read(mcb_engine *e, int kernel, double *total_ms, uint64_t *launches)
{
    if (!e || !total_ms || !launches) return fail(MCB_ERR_INVALID, "NULL argument");
    if (kernel < 0 || kernel >= MCB_KERNEL_COUNT) return fail(MCB_ERR_INVALID, "bad kernel id");
    DeviceGuard g(e->device);
    double ms = 0.0;
    uint64_t n = 0;
    std::vector<mcb_engine::TimedLaunch> keep;
    for (const auto &t : e->timed) {
        if (t.kernel != kernel) {
            keep.push_back(t);
            continue;
        }
        CU(cudaEventSynchronize(t.stop));
        float one = 0.0f;
        CU(cudaEventElapsedTime(&one, t.start, t.stop));
        ms += (double)one;
        ++n;
        e->event_pool.push_back(t);
    }
    e->timed.swap(keep);
    *total_ms = ms;
    *launches = n;
    return MCB_OK;
}

// -------------------------------------------------------------------------------- European
int mcb_european_segments_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                                int option_type, int rank, int world, double *d_segments, void *stream)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_shard(rank, world))) return rc;
    n_paths = resolve_paths(opt, n_paths);
    if (n_paths == 0) return fail(MCB_ERR_INVALID, "n_paths must be > 0");
    if (option_type != MCB_CALL && option_type != MCB_PUT) return fail(MCB_ERR_INVALID, "bad option_type");
    if (!d_segments) return fail(MCB_ERR_INVALID, "d_segments is NULL");
    DeviceGuard g(e->device);
    cudaStream_t st = pick(e, stream);
    const uint64_t n_chunks = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
    if ((rc = launch_european<MCB_EUROPEAN_PATHS_PER_SLOT>(e, prm, option_type, c_hi - c_lo, e->partials.ptr, nullptr,
                                                           0, st)))
        return rc;
    return launch_segments(e, e->partials.ptr, 0, c_lo, n_chunks, seg_lo, seg_hi, 1, d_segments, st);
}

int mcb_combine_segments_async(mcb_engine *e, const double *d_segments, int n_sets, uint64_t n_paths, float r, float T,
                               mcb_result *d_results, void *stream)
{
    if (!e || !d_segments || !d_results) return fail(MCB_ERR_INVALID, "NULL argument");
    if (n_sets < 1 || n_paths == 0) return fail(MCB_ERR_INVALID, "n_sets and n_paths must be positive");
    DeviceGuard g(e->device);
    const double discount = std::exp(-(double)r * (double)T);
    combine_kernel<<<(unsigned)n_sets, 32, 0, pick(e, stream)>>>(d_segments, n_paths, discount,
                                                                 reinterpret_cast<ResultDev *>(d_results));
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

static int finish_whole_job(mcb_engine *e, int n_sets, uint64_t n_paths, float r, float T, mcb_result *out)
{
    int rc;
    e->last_job_chunks = 0;   // mcb_last_segments: this job leaves its segments in h_segments
    if (n_sets == 1) {
        // single job (the reference's wrapper-sized calls): the result lands right behind the 64
        // segments, so ONE small D2H copy brings both back (a 1e6-path call is ~40 us end to end,
        // all of it launch + copy latency)
        static_assert(sizeof(mcb_result) == 5 * sizeof(double), "mcb_result is five 8-byte fields");
        double *d_job = e->segments.ptr;   // reserved to >= 2*MCB_SEGMENTS + 8 doubles at engine creation
        if ((rc = mcb_combine_segments_async(e, d_job, 1, n_paths, r, T,
                                             reinterpret_cast<mcb_result *>(d_job + 2 * MCB_SEGMENTS), nullptr)))
            return rc;
        CU(cudaMemcpyAsync(e->h_segments, d_job, sizeof(double) * (2 * MCB_SEGMENTS + 5), cudaMemcpyDeviceToHost,
                           e->stream));
        CU(cudaStreamSynchronize(e->stream));
        memcpy(out, e->h_segments + 2 * MCB_SEGMENTS, sizeof(mcb_result));
        return MCB_OK;
    }
    if ((rc = reserve_results(e, (size_t)n_sets))) return rc;
    if ((rc = mcb_combine_segments_async(e, e->segments.ptr, n_sets, n_paths, r, T,
                                         reinterpret_cast<mcb_result *>(e->results.ptr), nullptr)))
        return rc;
    CU(cudaMemcpyAsync(e->h_results, e->results.ptr, sizeof(mcb_result) * (size_t)n_sets, cudaMemcpyDeviceToHost,
                       e->stream));
    CU(cudaMemcpyAsync(e->h_segments, e->segments.ptr, sizeof(double) * 2 * MCB_SEGMENTS, cudaMemcpyDeviceToHost,
                       e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memcpy(out, e->h_results, sizeof(mcb_result) * (size_t)n_sets);
    return MCB_OK;
}

int mcb_price_european(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int option_type,
                       mcb_result *out)
{
    // ONE launch per shard; the result comes back through mapped host memory (no copy, no stream sync)
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    uint64_t ticket = 0;
    int rc = mcb_european_submit(e, opt, n_paths, seed, option_type, &ticket);
    if (rc) return rc;
    return mcb_european_collect(e, ticket, out);
}

// Whole-job calls on a multi-device engine: every shard fills the segments it owns straight into the
// LEADER's segment buffer (peer stores over NVLink from inside segment_kernel), the leader's stream
// waits for the shards' events and runs the fixed final tree.  No collective library, same bits.
// (The whole-job calls are synchronous, so the leader's buffer is free when the next one starts.)
extern "C++" {
template <typename F>
static int run_on_shards(mcb_engine *e, F &&enqueue)
{
    const size_t n = shard_count(e);
    // every shard's launcher thread enqueues its own device (parameter staging + launches + its event) while this
    // thread does the leader's; then the leader's stream waits for the shards' events
    auto one = [&](mcb_engine *s, size_t i) -> int {
        DeviceGuard g(s->device);
        int rc = enqueue(s, (int)i, (int)n);
        if (rc == MCB_OK && i) {
            cudaError_t err = cudaEventRecord(s->p_done[0], s->stream);
            if (err != cudaSuccess) rc = fail(MCB_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(err));
        }
        return rc;
    };
    for (size_t i = 1; i < n; ++i) {
        mcb_engine *s = shard_at(e, i);
        if (!s->worker) continue;
        s->worker->post([&one, s, i]() {
            const int r = one(s, i);
            if (r != MCB_OK) s->worker->message = g_error;
            return r;
        });
    }
    int rc = one(e, 0);
    for (size_t i = 1; i < n; ++i) {
        mcb_engine *s = shard_at(e, i);
        const int r = s->worker ? s->worker->wait() : one(s, i);
        if (r != MCB_OK && rc == MCB_OK)
            rc = s->worker ? fail(r, "shard %zu (device %d): %s", i, s->device, s->worker->message.c_str()) : r;
    }
    if (rc) return rc;
    DeviceGuard g(e->device);
    for (size_t i = 1; i < n; ++i) CU(cudaStreamWaitEvent(e->stream, shard_at(e, i)->p_done[0], 0));
    return MCB_OK;
}
}  // extern "C++"

// Slab-sharded modes (trajectories, nested MC) on a multi-device engine: one host thread per shard
// drives that shard's own synchronous call on its slab of paths; rows are pure functions of
// (seed, path id), so the slabs concatenate to the single-device result.  No collective.
extern "C++" {
template <typename F>
static int threads_over_shards(mcb_engine *e, F &&body)
{
    const size_t n = shard_count(e);
    std::vector<mcb_engine *> list;
    for (size_t i = 0; i < n; ++i) list.push_back(shard_at(e, i));
    std::vector<int> rcs(n, MCB_OK);
    std::vector<std::string> msgs(n);
    std::vector<std::thread> threads;
    for (size_t i = 0; i < n; ++i)
        threads.emplace_back([&, i]() {
            rcs[i] = body(list[i], (int)i, (int)n);
            if (rcs[i] != MCB_OK) msgs[i] = g_error;   // g_error is thread-local: carry the message over
        });
    for (auto &t : threads) t.join();
    for (size_t i = 0; i < n; ++i)
        if (rcs[i] != MCB_OK) return fail(rcs[i], "shard %zu (device %d): %s", i, list[i]->device, msgs[i].c_str());
    return MCB_OK;
}
}  // extern "C++"

// ------------------------------------------------------------------------- European, packed keying
static int european_packed_segments_impl(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                                         int option_type, int rank, int world, double *d_segments, int write_unowned)
{
    DeviceGuard g(e->device);
    cudaStream_t st = e->stream;
    const uint64_t n_chunks = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    if (c_hi - c_lo > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    int rc;
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
    if (c_hi > c_lo) {
        {
            TimedScope timed(e, MCB_KERNEL_EUROPEAN_PACKED, st);
            if (option_type == MCB_PUT)
                european_packed_kernel<kPut, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(
                    prm, e->partials.ptr, nullptr, 0);
            else
                european_packed_kernel<kCall, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(
                    prm, e->partials.ptr, nullptr, 0);
        }
        e->launches++;
        CU(cudaGetLastError());
    }
    return launch_segments(e, e->partials.ptr, 0, c_lo, n_chunks, seg_lo, seg_hi, 1, d_segments, st, write_unowned);
}

int mcb_price_european_packed(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                              int option_type, mcb_result *out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (e->leader) return fail(MCB_ERR_INVALID, "call the multi-device engine, not one of its shards");
    if (option_type != MCB_CALL && option_type != MCB_PUT) return fail(MCB_ERR_INVALID, "bad option_type");
    n_paths = resolve_paths(opt, n_paths);
    if (n_paths == 0) return fail(MCB_ERR_INVALID, "n_paths must be > 0");
    DeviceGuard g(e->device);
    double *dst = e->segments.ptr;
    if ((rc = run_on_shards(e, [&](mcb_engine *s, int rank, int world) {
             return european_packed_segments_impl(s, opt, n_paths, seed, option_type, rank, world, dst, world == 1);
         })))
        return rc;
    return finish_whole_job(e, 1, n_paths, opt->r, opt->T, out);
}

int mcb_european_packed_payoffs(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                                uint64_t seed, int option_type, float *payoffs)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (!payoffs) return fail(MCB_ERR_INVALID, "payoffs is NULL");
    if (n_paths == 0) return MCB_OK;
    DeviceGuard g(e->device);
    const uint64_t c_lo = first_path / kEuropeanChunk;
    const uint64_t c_hi = (first_path + n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    if (c_hi - c_lo > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    if ((rc = e->scratch.reserve((size_t)n_paths * sizeof(float)))) return rc;
    float *d = reinterpret_cast<float *>(e->scratch.ptr);
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, first_path + n_paths, seed, c_lo);
    if (option_type == MCB_PUT)
        european_packed_kernel<kPut, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, e->stream>>>(
            prm, e->partials.ptr, d, first_path);
    else
        european_packed_kernel<kCall, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, e->stream>>>(
            prm, e->partials.ptr, d, first_path);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(payoffs, d, (size_t)n_paths * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

// ---------------------------------------------------------------------------------- bullet
static int bullet_params(const mcb_option_data *opt, uint64_t n_paths_end, uint64_t seed, int Ik, float Sk, int Tk,
                         uint64_t first_chunk, WalkParams *out)
{
    int rc = check_walk(opt);
    if (rc) return rc;
    if (Tk < 0 || Tk > opt->N_STEPS) return fail(MCB_ERR_INVALID, "Tk must be in [0, N_STEPS]");
    if (Sk < 0.0f || !std::isfinite(Sk)) return fail(MCB_ERR_INVALID, "Sk must be >= 0 and finite");
    // Sk == 0 means "start from S0" exactly as inc/trajectories.cuh:141
    const WalkConsts w = walk_consts(opt, Sk == 0.0f ? (double)opt->S0 : (double)Sk);
    WalkParams p{};
    const bool barrier = std::isfinite(w.lBd);
    p.n_steps = opt->N_STEPS - Tk;
    p.sc = w.sc;
    // constants of the drift-free walk (pricing_kernels.cuh), formed in double from the SAME float sc / dr the
    // trajectory kernels use, rounded once
    p.acc0 = barrier ? (float)(-(w.lBd - w.l0d) / (double)w.sc) : 0.0f;
    p.bq = barrier ? (float)((double)w.dr / (double)w.sc) : INFINITY;
    p.l_end = (float)((barrier ? w.lBd : w.l0d) + (double)p.n_steps * (double)w.dr);
    p.K = opt->K; p.P1 = opt->P1; p.P2 = opt->P2;
    p.count0 = Ik;
    p.table = p.n_steps <= kWalkTableMaxSteps ? 1 : 0;
    p.n_paths = n_paths_end;
    p.first_chunk = first_chunk;
    p.keys = make_philox_keys(seed);
    *out = p;
    return MCB_OK;
}

static size_t walk_table_bytes(const WalkParams &p)
{
    return p.table ? (size_t)((p.n_steps + 3) & ~3) * sizeof(float) : 0;
}

static int bullet_segments_impl(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int Ik,
                                float Sk, int Tk, int rank, int world, double *d_segments, void *stream,
                                int write_unowned)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_shard(rank, world))) return rc;
    n_paths = resolve_paths(opt, n_paths);
    if (n_paths == 0) return fail(MCB_ERR_INVALID, "n_paths must be > 0");
    if (!d_segments) return fail(MCB_ERR_INVALID, "d_segments is NULL");
    DeviceGuard g(e->device);
    cudaStream_t st = pick(e, stream);
    const uint64_t n_chunks = (n_paths + kBulletChunk - 1) / kBulletChunk;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    if (c_hi - c_lo > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    WalkParams prm;
    if ((rc = bullet_params(opt, n_paths, seed, Ik, Sk, Tk, c_lo, &prm))) return rc;
    if (c_hi > c_lo) {
        {
            TimedScope timed(e, MCB_KERNEL_BULLET, st);
            bullet_kernel<MCB_BULLET_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, walk_table_bytes(prm), st>>>(
                prm, e->partials.ptr, nullptr, 0);
        }
        e->launches++;
        CU(cudaGetLastError());
    }
    return launch_segments(e, e->partials.ptr, 0, c_lo, n_chunks, seg_lo, seg_hi, 1, d_segments, st, write_unowned);
}

int mcb_bullet_segments_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int Ik,
                              float Sk, int Tk, int rank, int world, double *d_segments, void *stream)
{
    return bullet_segments_impl(e, opt, n_paths, seed, Ik, Sk, Tk, rank, world, d_segments, stream, 1);
}

int mcb_price_bullet(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int Ik, float Sk,
                     int Tk, mcb_result *out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    int rc = check_common(e, opt);
    if (rc) return rc;
    DeviceGuard g(e->device);
    n_paths = resolve_paths(opt, n_paths);
    double *dst = e->segments.ptr;
    if ((rc = run_on_shards(e, [&](mcb_engine *s, int rank, int world) {
             return bullet_segments_impl(s, opt, n_paths, seed, Ik, Sk, Tk, rank, world, dst, nullptr, world == 1);
         })))
        return rc;
    return finish_whole_job(e, 1, n_paths, opt->r, opt->T, out);
}

// ------------------------------------------------------------------------------------ sweep
static int sweep_segments_impl(mcb_engine *e, const mcb_option_data *opt, const float *strikes, const float *vols,
                               int n_params, uint64_t n_paths, uint64_t seed, int option_type, int rank, int world,
                               double *d_segments, void *stream, int write_unowned)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_shard(rank, world))) return rc;
    if (!strikes || !vols || n_params < 1) return fail(MCB_ERR_INVALID, "bad parameter arrays");
    n_paths = resolve_paths(opt, n_paths);
    if (n_paths == 0) return fail(MCB_ERR_INVALID, "n_paths must be > 0");
    if (option_type != MCB_CALL && option_type != MCB_PUT) return fail(MCB_ERR_INVALID, "bad option_type");
    if (!d_segments) return fail(MCB_ERR_INVALID, "d_segments is NULL");
    for (int i = 0; i < n_params; ++i)
        if (!std::isfinite(strikes[i]) || !(vols[i] >= 0.0f) || !std::isfinite(vols[i]))
            return fail(MCB_ERR_INVALID, "parameter set %d is not finite / has negative vol", i);
    DeviceGuard g(e->device);
    cudaStream_t st = pick(e, stream);
    const uint64_t n_chunks = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    const uint64_t local = c_hi - c_lo;
    if (local > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    // Parameter sets are priced in groups so the partials workspace stays bounded (<= 64 MiB);
    // within a group ONE launch draws every path once and walks all the sets (sweep_kernel).
    uint64_t group = local ? (uint64_t)(8u << 20) / local : (uint64_t)n_params;
    if (group < 1) group = 1;
    if (group > 65535) group = 65535;   // segment_kernel puts the parameter set on gridDim.y
    if (group > (uint64_t)n_params) group = (uint64_t)n_params;
    const uint64_t stride = local + 1;
    if ((rc = e->partials.reserve((size_t)(group * stride)))) return rc;
    // per-set constants (c0, c1, K) folded in double exactly as european_params does
    if ((rc = e->sweep_sets.reserve((size_t)n_params))) return rc;
    // staged from pageable memory: cudaMemcpyAsync returns once the source has been read, so the
    // vector can be refilled by the next call without any event bookkeeping
    e->h_sweep_sets.resize((size_t)n_params);
    for (int i = 0; i < n_params; ++i) {
        const EuropeanParams one = european_params(opt, strikes[i], vols[i], n_paths, seed, c_lo);
        e->h_sweep_sets[i] = make_float4(one.c0, one.c1, one.K, 0.0f);
    }
    CU(cudaMemcpyAsync(e->sweep_sets.ptr, e->h_sweep_sets.data(), sizeof(float4) * (size_t)n_params,
                       cudaMemcpyHostToDevice, st));
    SweepParams prm{};
    prm.n_paths = n_paths;
    prm.first_chunk = c_lo;
    prm.stride = stride;
    prm.keys = make_philox_keys(seed);
    for (uint64_t i0 = 0; i0 < (uint64_t)n_params; i0 += group) {
        const uint64_t cnt = (uint64_t)n_params - i0 < group ? (uint64_t)n_params - i0 : group;
        prm.n_sets = (int)cnt;
        // How many ways to split the sets over gridDim.y: a CTA costs one draw of its chunk (~57 instructions
        // per path, about 10.5 (path, set) units of the MUFU-bound loop) plus its sets, and the grid runs in
        // whole waves of 2 CTAs per SM (128 registers).  More splits repeat the draw but fill the last wave:
        // 512 chunks x 1024 sets on 148 SMs -> 4 splits (6.92 waves of 256-set CTAs) instead of 2 (3.46 waves).
        const uint64_t slots = 2ull * (uint64_t)e->prop.multiProcessorCount;
        uint64_t best_sets = (cnt + kSweepTile - 1) / kSweepTile * kSweepTile;
        double best_cost = 1e300;
        for (uint64_t y : {1ull, 2ull, 3ull, 4ull, 6ull, 8ull}) {
            if (y > 1 && cnt / y < 64) break;
            const uint64_t per = ((cnt + y - 1) / y + kSweepTile - 1) / kSweepTile * kSweepTile;
            const uint64_t gy = (cnt + per - 1) / per;
            const uint64_t waves = local ? (local * gy + slots - 1) / slots : 0;
            const double cost = (double)waves * (10.5 + (double)per);
            if (cost < best_cost * 0.995) {
                best_cost = cost;
                best_sets = per;
            }
        }
        prm.sets_per_cta = (int)best_sets;
        const unsigned grid_y = (unsigned)((cnt + (uint64_t)prm.sets_per_cta - 1) / (uint64_t)prm.sets_per_cta);
        if (local) {
            TimedScope timed(e, MCB_KERNEL_SWEEP, st);
            const dim3 grid((unsigned)local, grid_y);
            if (option_type == MCB_PUT)
                sweep_kernel<kPut, MCB_EUROPEAN_PATHS_PER_SLOT><<<grid, kSlots, 0, st>>>(
                    prm, e->sweep_sets.ptr + i0, e->partials.ptr);
            else
                sweep_kernel<kCall, MCB_EUROPEAN_PATHS_PER_SLOT><<<grid, kSlots, 0, st>>>(
                    prm, e->sweep_sets.ptr + i0, e->partials.ptr);
            e->launches++;
            CU(cudaGetLastError());
        }
        if ((rc = launch_segments(e, e->partials.ptr, stride, c_lo, n_chunks, seg_lo, seg_hi, (int)cnt,
                                  d_segments + i0 * 2 * MCB_SEGMENTS, st, write_unowned)))
            return rc;
    }
    return MCB_OK;
}

int mcb_sweep_segments_async(mcb_engine *e, const mcb_option_data *opt, const float *strikes, const float *vols,
                             int n_params, uint64_t n_paths, uint64_t seed, int option_type, int rank, int world,
                             double *d_segments, void *stream)
{
    return sweep_segments_impl(e, opt, strikes, vols, n_params, n_paths, seed, option_type, rank, world, d_segments,
                               stream, 1);
}

int mcb_price_sweep(mcb_engine *e, const mcb_option_data *opt, const float *strikes, const float *vols, int n_params,
                    uint64_t n_paths, uint64_t seed, int option_type, mcb_result *out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (n_params < 1) return fail(MCB_ERR_INVALID, "n_params must be >= 1");
    DeviceGuard g(e->device);
    n_paths = resolve_paths(opt, n_paths);
    if ((rc = e->segments.reserve((size_t)n_params * 2 * MCB_SEGMENTS))) return rc;
    double *dst = e->segments.ptr;
    if ((rc = run_on_shards(e, [&](mcb_engine *s, int rank, int world) {
             return sweep_segments_impl(s, opt, strikes, vols, n_params, n_paths, seed, option_type, rank, world, dst,
                                        nullptr, world == 1);
         })))
        return rc;
    return finish_whole_job(e, n_params, n_paths, opt->r, opt->T, out);
}

// ------------------------------------------------------------------------------ trajectories
static int trajectories_launch(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                               uint64_t seed, float *d_prices, int *d_counts, float *d_logs, void *stream)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_walk(opt))) return rc;
    if (!d_prices) return fail(MCB_ERR_INVALID, "prices is NULL");
    if (n_paths == 0) return MCB_OK;
    DeviceGuard g(e->device);
    const WalkConsts w = walk_consts(opt, (double)opt->S0);
    PathParams prm{};
    prm.l0 = w.l0; prm.sc = w.sc; prm.dr = w.dr; prm.lB = w.lB;
    prm.n_steps = opt->N_STEPS;
    prm.first_path = first_path;
    prm.n_paths = n_paths;
    prm.keys = make_philox_keys(seed);
    if (n_paths > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many paths for one launch");
    const bool base_aligned = ((uintptr_t)d_prices % 16 == 0) && (!d_counts || (uintptr_t)d_counts % 16 == 0) &&
                              (!d_logs || (uintptr_t)d_logs % 16 == 0);
    const bool vec = (opt->N_STEPS % 4 == 0) && base_aligned;
    cudaStream_t st = pick(e, stream);
    {
        TimedScope timed(e, MCB_KERNEL_TRAJECTORY, st);
        // Row layout (steps per lane x lanes per row), a function of n_steps ONLY so that a row's
        // bits never depend on which arrays were asked for or which kernel wrote it: the smallest
        // pass that holds the whole row, 16 x 16 (several passes) beyond 256 steps.
        const int n = opt->N_STEPS;
        if (n <= 32) launch_trajectory<4, 8>(prm, n_paths, vec, base_aligned, d_prices, d_counts, d_logs, st);
        else if (n <= 64) launch_trajectory<4, 16>(prm, n_paths, vec, base_aligned, d_prices, d_counts, d_logs, st);
        else if (n <= 128) launch_trajectory<8, 16>(prm, n_paths, vec, base_aligned, d_prices, d_counts, d_logs, st);
        else if (n <= 192) launch_trajectory<12, 16>(prm, n_paths, vec, base_aligned, d_prices, d_counts, d_logs, st);
        else launch_trajectory<16, 16>(prm, n_paths, vec, base_aligned, d_prices, d_counts, d_logs, st);
    }
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

int mcb_trajectories_async(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                           uint64_t seed, float *d_prices, int *d_counts, void *stream)
{
    return trajectories_launch(e, opt, first_path, n_paths, seed, d_prices, d_counts, nullptr, stream);
}

static int trajectories_to_host(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                                uint64_t seed, float *prices, int *counts);

int mcb_simulate_trajectories(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                              uint64_t seed, float *prices, int *counts, int where)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_walk(opt))) return rc;
    if (!prices) return fail(MCB_ERR_INVALID, "prices is NULL");
    DeviceGuard g(e->device);
    if (where == MCB_DEVICE) {
        if ((rc = mcb_trajectories_async(e, opt, first_path, n_paths, seed, prices, counts, nullptr))) return rc;
        CU(cudaStreamSynchronize(e->stream));
        return MCB_OK;
    }
    if (where != MCB_HOST) return fail(MCB_ERR_INVALID, "bad `where`");
    if (n_paths == 0) return MCB_OK;
    if (shard_count(e) > 1) {   // multi-device engine: contiguous slabs, every shard copies its own rows home
        const size_t row = (size_t)opt->N_STEPS;
        return threads_over_shards(e, [&](mcb_engine *s, int rank, int world) {
            const uint64_t lo = n_paths * (uint64_t)rank / (uint64_t)world, hi = n_paths * (uint64_t)(rank + 1) / (uint64_t)world;
            if (hi == lo) return (int)MCB_OK;
            return trajectories_to_host(s, opt, first_path + lo, hi - lo, seed, prices + lo * row,
                                        counts ? counts + lo * row : nullptr);
        });
    }
    return trajectories_to_host(e, opt, first_path, n_paths, seed, prices, counts);
}

static int trajectories_to_host(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                                uint64_t seed, float *prices, int *counts)
{
    int rc;
    DeviceGuard g(e->device);
    // Host destination: rows are pure functions of (seed, path id), so the job is cut into slabs of
    // <= 128 MB that go through the engine's grow-only workspace (no per-call cudaMalloc / cudaFree,
    // bounded device memory whatever n_paths is) and are copied back slab by slab.
    const size_t row = (size_t)opt->N_STEPS;
    uint64_t slab_rows = ((size_t)32 << 20) / row;   // 32 Mi floats = 128 MB per array
    if (slab_rows < 1) slab_rows = 1;
    if (slab_rows > n_paths) slab_rows = n_paths;
    const size_t slab_elems = (((size_t)slab_rows * row) + 3) & ~(size_t)3;   // keeps the second array 16-byte aligned
    if ((rc = e->traj_ws.reserve(slab_elems * (counts ? 2 : 1)))) return rc;
    float *dp = e->traj_ws.ptr;
    int *dc = counts ? reinterpret_cast<int *>(e->traj_ws.ptr + slab_elems) : nullptr;
    for (uint64_t done = 0; done < n_paths; done += slab_rows) {
        const uint64_t rows = n_paths - done < slab_rows ? n_paths - done : slab_rows;
        if ((rc = mcb_trajectories_async(e, opt, first_path + done, rows, seed, dp, dc, nullptr))) return rc;
        const size_t off = (size_t)done * row, cnt = (size_t)rows * row;
        CU(cudaMemcpyAsync(prices + off, dp, cnt * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        if (counts) CU(cudaMemcpyAsync(counts + off, dc, cnt * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));   // the workspace is reused by the next slab
    }
    return MCB_OK;
}

// ---------------------------------------------------------------------------------- nested MC
int mcb_nested_async(mcb_engine *e, const mcb_option_data *opt, uint64_t first_outer, uint64_t n_outer,
                     uint64_t seed_outer, uint64_t seed_inner, int discount_mode, float *d_F, float *d_prices,
                     int *d_counts, void *stream)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_walk(opt))) return rc;
    if (!d_F) return fail(MCB_ERR_INVALID, "F is NULL");
    if (opt->N_PATHS_INNER < 1) return fail(MCB_ERR_INVALID, "N_PATHS_INNER must be >= 1");
    if (discount_mode != MCB_DISCOUNT_COMPAT && discount_mode != MCB_DISCOUNT_CORRECT)
        return fail(MCB_ERR_INVALID, "bad discount_mode");
    if (n_outer == 0) return MCB_OK;
    if (n_outer > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many outer paths for one launch");
    DeviceGuard g(e->device);
    // outer walk: trajectory_kernel, which also leaves log2 S and the barrier count of every
    // point in the engine's workspace (prices / counts go to the caller's buffers when given)
    const size_t n = (size_t)n_outer * (size_t)opt->N_STEPS;
    const size_t n4 = (n + 3) & ~(size_t)3;  // keeps every sub-buffer 16-byte aligned
    if ((rc = e->nested_ws.reserve(n4 * 3))) return rc;
    float *ws_logs = e->nested_ws.ptr;
    float *ws_prices = d_prices ? d_prices : e->nested_ws.ptr + n4;
    int *ws_counts = d_counts ? d_counts : reinterpret_cast<int *>(e->nested_ws.ptr + 2 * n4);
    if ((rc = trajectories_launch(e, opt, first_outer, n_outer, seed_outer, ws_prices, ws_counts, ws_logs, stream)))
        return rc;
    const WalkConsts w = walk_consts(opt, (double)opt->S0);
    if (opt->N_STEPS > kWalkTableMaxSteps)
        return fail(MCB_ERR_INVALID, "nested MC supports at most %d steps", kWalkTableMaxSteps);
    NestedParams prm{};
    prm.sc = w.sc; prm.dr = w.dr; prm.lB = w.lB;
    prm.inv_sc = (float)(1.0 / (double)w.sc);
    prm.bq = std::isfinite(w.lBd) ? (float)((double)w.dr / (double)w.sc) : INFINITY;
    prm.K = opt->K; prm.P1 = opt->P1; prm.P2 = opt->P2;
    prm.n_steps = opt->N_STEPS;
    prm.n_inner = opt->N_PATHS_INNER;
    prm.discount_mode = discount_mode;
    prm.r = opt->r; prm.T = opt->T; prm.dt = opt->step;
    prm.first_outer = first_outer;
    prm.keys_inner = make_philox_keys(seed_inner);
    {
        cudaStream_t st = pick(e, stream);
        TimedScope timed(e, MCB_KERNEL_NESTED, st);
        // points of one outer trajectory are spread over `split` CTAs (interleaved k): at least 8 (C4:
        // 56.9 -> 54.2 ms), more when there are few outer trajectories, so the grid stays at >= ~8 waves
        const uint64_t want = 8ull * 5ull * (uint64_t)e->prop.multiProcessorCount;
        uint64_t split = (want + n_outer - 1) / n_outer;
        if (split < 8) split = 8;
        if (split > (uint64_t)opt->N_STEPS) split = (uint64_t)opt->N_STEPS;
        const size_t thr_bytes = (size_t)((opt->N_STEPS + 3) & ~3) * sizeof(float);
        nested_kernel<<<dim3((unsigned)n_outer, (unsigned)split), kSlots, thr_bytes, st>>>(prm, ws_logs, ws_counts, d_F);
    }
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

static int nested_single(mcb_engine *e, const mcb_option_data *opt, uint64_t first_outer, uint64_t n_outer,
                         uint64_t seed_outer, uint64_t seed_inner, int discount_mode, float *F, float *prices,
                         int *counts, int where, double *mean_F);

int mcb_nested_monte_carlo(mcb_engine *e, const mcb_option_data *opt, uint64_t first_outer, uint64_t n_outer,
                           uint64_t seed_outer, uint64_t seed_inner, int discount_mode, float *F, float *prices,
                           int *counts, int where, double *mean_F)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if ((rc = check_walk(opt))) return rc;
    if (!F) return fail(MCB_ERR_INVALID, "F is NULL");
    if (where != MCB_HOST && where != MCB_DEVICE) return fail(MCB_ERR_INVALID, "bad `where`");
    DeviceGuard g(e->device);
    const size_t n = (size_t)n_outer * (size_t)opt->N_STEPS;
    if (n == 0) {
        if (mean_F) *mean_F = 0.0;
        return MCB_OK;
    }
    if (shard_count(e) > 1 && where == MCB_HOST) {
        // multi-device engine: outer trajectory p and everything hanging off it depend on p only
        const size_t row = (size_t)opt->N_STEPS;
        rc = threads_over_shards(e, [&](mcb_engine *s, int rank, int world) {
            const uint64_t lo = n_outer * (uint64_t)rank / (uint64_t)world, hi = n_outer * (uint64_t)(rank + 1) / (uint64_t)world;
            if (hi == lo) return (int)MCB_OK;
            return nested_single(s, opt, first_outer + lo, hi - lo, seed_outer, seed_inner, discount_mode,
                                 F + lo * row, prices ? prices + lo * row : nullptr,
                                 counts ? counts + lo * row : nullptr, MCB_HOST, nullptr);
        });
        if (rc) return rc;
        if (mean_F) {
            double acc = 0.0;
            for (size_t i = 0; i < n; ++i) acc += (double)F[i];
            *mean_F = acc / (double)(n + 1);
        }
        return MCB_OK;
    }
    return nested_single(e, opt, first_outer, n_outer, seed_outer, seed_inner, discount_mode, F, prices, counts, where,
                         mean_F);
}

static int nested_single(mcb_engine *e, const mcb_option_data *opt, uint64_t first_outer, uint64_t n_outer,
                         uint64_t seed_outer, uint64_t seed_inner, int discount_mode, float *F, float *prices,
                         int *counts, int where, double *mean_F)
{
    int rc;
    DeviceGuard g(e->device);
    const size_t n = (size_t)n_outer * (size_t)opt->N_STEPS;
    float *dF = F, *dP = prices;
    int *dC = counts;
    std::vector<float> hostF;
    if (where == MCB_HOST) {
        const size_t bytes = n * sizeof(float) * (prices ? 2 : 1) + (counts ? n * sizeof(int) : 0);
        if ((rc = e->scratch.reserve(bytes))) return rc;
        dF = reinterpret_cast<float *>(e->scratch.ptr);
        dP = prices ? dF + n : nullptr;
        dC = counts ? reinterpret_cast<int *>(dF + (prices ? 2 * n : n)) : nullptr;
    }
    if ((rc = mcb_nested_async(e, opt, first_outer, n_outer, seed_outer, seed_inner, discount_mode, dF, dP, dC,
                               nullptr)))
        return rc;
    const float *sumsrc = nullptr;
    if (where == MCB_HOST) {
        CU(cudaMemcpyAsync(F, dF, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        if (prices) CU(cudaMemcpyAsync(prices, dP, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        if (counts) CU(cudaMemcpyAsync(counts, dC, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        sumsrc = F;
    } else {
        if (mean_F) {
            hostF.resize(n);
            CU(cudaMemcpyAsync(hostF.data(), dF, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        }
        CU(cudaStreamSynchronize(e->stream));
        sumsrc = hostF.data();
    }
    if (mean_F) {
        // the wrappers' diagnostic scalar: mean over N*steps + 1 slots (inc/wrappers.cuh:134,185-189)
        double s = 0.0;
        for (size_t i = 0; i < n; ++i) s += (double)sumsrc[i];
        *mean_F = s / (double)(n + 1);
    }
    return MCB_OK;
}

// ------------------------------------------------------------------- reduce / pre-generated
int mcb_reduce_sum(mcb_engine *e, const float *x, uint64_t n, int where, float *out)
{
    if (!e || !out || (!x && n)) return fail(MCB_ERR_INVALID, "NULL argument");
    if (where != MCB_HOST && where != MCB_DEVICE) return fail(MCB_ERR_INVALID, "bad `where`");
    DeviceGuard g(e->device);
    int rc;
    const float *dx = x;
    if (where == MCB_HOST) {
        if ((rc = e->scratch.reserve((size_t)n * sizeof(float) + 16))) return rc;
        if (n) CU(cudaMemcpyAsync(e->scratch.ptr, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        dx = reinterpret_cast<const float *>(e->scratch.ptr);
    }
    if ((rc = reserve_results(e, 1))) return rc;
    float *d_out = reinterpret_cast<float *>(e->results.ptr);
    reduce_sum_kernel<<<1, kSlots, 0, e->stream>>>(dx, n, d_out);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(e->h_results, d_out, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memcpy(out, e->h_results, sizeof(float));
    return MCB_OK;
}

int mcb_reduce_blocks(mcb_engine *e, const float *x, uint64_t n, int where, uint32_t n_blocks, uint64_t span,
                      int strided, float *out)
{
    if (!e || !out || (!x && n)) return fail(MCB_ERR_INVALID, "NULL argument");
    if (where != MCB_HOST && where != MCB_DEVICE) return fail(MCB_ERR_INVALID, "bad `where`");
    if (n_blocks == 0 || n_blocks > 0x7fffffffu || span == 0) return fail(MCB_ERR_INVALID, "bad n_blocks / span");
    DeviceGuard g(e->device);
    int rc;
    const size_t in_bytes = where == MCB_HOST ? (size_t)n * sizeof(float) : 0;
    const size_t in_pad = (in_bytes + 255) & ~(size_t)255;
    if ((rc = e->scratch.reserve(in_pad + (size_t)n_blocks * sizeof(float) + 16))) return rc;
    const float *dx = x;
    if (where == MCB_HOST) {
        if (n) CU(cudaMemcpyAsync(e->scratch.ptr, x, in_bytes, cudaMemcpyHostToDevice, e->stream));
        dx = reinterpret_cast<const float *>(e->scratch.ptr);
    }
    float *d_out = reinterpret_cast<float *>(e->scratch.ptr + in_pad);
    reduce_blocks_kernel<<<n_blocks, kSlots, 0, e->stream>>>(dx, n, span, strided, d_out);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, (size_t)n_blocks * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_generate_normals(mcb_engine *e, uint64_t seed, uint64_t n, float *out, int where)
{
    if (!e || (!out && n)) return fail(MCB_ERR_INVALID, "NULL argument");
    if (where != MCB_HOST && where != MCB_DEVICE) return fail(MCB_ERR_INVALID, "bad `where`");
    if (n == 0) return MCB_OK;
    if (where == MCB_HOST) return mcb_stream_normals(e, seed, 0, 0, n, out);
    DeviceGuard g(e->device);
    if ((n + 127) / 128 > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many normals for one launch");
    stream_normals_kernel<<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(make_philox_keys(seed), 0, 0, n, out);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_write_trajectories_csv(const char *path, const float *prices, uint64_t n_trajectories, int n_steps, float x0,
                               float dt)
{
    if (!path || (!prices && n_trajectories)) return fail(MCB_ERR_INVALID, "NULL argument");
    if (n_steps < 1) return fail(MCB_ERR_INVALID, "n_steps must be >= 1");
    FILE *f = fopen(path, "w");
    if (!f) return fail(MCB_ERR_INVALID, "cannot open %s for writing", path);
    // same text a C++ ostream prints for these floats (testing.cu:41-46 streams float values)
    fprintf(f, "time,trajectory,value\n");
    for (uint64_t p = 0; p < n_trajectories; ++p) {
        fprintf(f, "0,%llu,%g\n", (unsigned long long)p, (double)x0);
        for (int i = 0; i < n_steps; ++i)
            fprintf(f, "%g,%llu,%g\n", (double)((float)(1 + i) * dt), (unsigned long long)p,
                    (double)prices[p * (uint64_t)n_steps + (uint64_t)i]);
    }
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) return fail(MCB_ERR_INVALID, "write to %s failed", path);
    return MCB_OK;
}

int mcb_price_from_normals(mcb_engine *e, const mcb_option_data *opt, const float *normals, uint64_t n_paths,
                           int n_steps, float *payoffs, int where)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (!normals || !payoffs) return fail(MCB_ERR_INVALID, "NULL argument");
    if (n_steps < 1 || !(opt->step > 0.0f) || !(opt->v > 0.0f)) return fail(MCB_ERR_INVALID, "bad steps/dt/sigma");
    if (where != MCB_HOST && where != MCB_DEVICE) return fail(MCB_ERR_INVALID, "bad `where`");
    if (n_paths == 0) return MCB_OK;
    DeviceGuard g(e->device);
    const WalkConsts w = walk_consts(opt, (double)opt->S0);
    const size_t nz = (size_t)n_paths * (size_t)n_steps;
    const float *dz = normals;
    float *dp = payoffs;
    if (where == MCB_HOST) {
        if ((rc = e->scratch.reserve((nz + (size_t)n_paths) * sizeof(float)))) return rc;
        float *base = reinterpret_cast<float *>(e->scratch.ptr);
        CU(cudaMemcpyAsync(base, normals, nz * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        dz = base;
        dp = base + nz;
    }
    const uint64_t ctas = (n_paths + kWarps - 1) / kWarps;   // one warp per path
    if (ctas > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many paths for one launch");
    pregen_kernel<<<(unsigned)ctas, kSlots, 0, e->stream>>>(dz, n_paths, n_steps, w.l0, w.dr, w.v, opt->K, dp);
    e->launches++;
    CU(cudaGetLastError());
    if (where == MCB_HOST)
        CU(cudaMemcpyAsync(payoffs, dp, (size_t)n_paths * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

// ------------------------------------------------------------------------------ parity hooks
static int blocks_hook(mcb_engine *e, uint64_t seed, const uint64_t *subsequences, const uint64_t *blocks, uint64_t n,
                       uint32_t *words, bool library)
{
    if (!e || !subsequences || !blocks || !words) return fail(MCB_ERR_INVALID, "NULL argument");
    if (n == 0) return MCB_OK;
    DeviceGuard g(e->device);
    int rc;
    const size_t in_bytes = (size_t)n * sizeof(uint64_t);
    if ((rc = e->scratch.reserve(2 * in_bytes + (size_t)n * 16))) return rc;
    uint64_t *d_sub = reinterpret_cast<uint64_t *>(e->scratch.ptr);
    uint64_t *d_blk = d_sub + n;
    uint4 *d_out = reinterpret_cast<uint4 *>(d_blk + n);
    CU(cudaMemcpyAsync(d_sub, subsequences, in_bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_blk, blocks, in_bytes, cudaMemcpyHostToDevice, e->stream));
    const unsigned ctas = (unsigned)((n + 127) / 128);
    if (library)
        curand_blocks_kernel<<<ctas, 128, 0, e->stream>>>(seed, d_sub, d_blk, n, d_out);
    else
        philox_blocks_kernel<<<ctas, 128, 0, e->stream>>>(make_philox_keys(seed), d_sub, d_blk, n, d_out);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(words, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_philox_blocks(mcb_engine *e, uint64_t seed, const uint64_t *subsequences, const uint64_t *blocks, uint64_t n,
                      uint32_t *words)
{
    return blocks_hook(e, seed, subsequences, blocks, n, words, false);
}

int mcb_curand_blocks(mcb_engine *e, uint64_t seed, const uint64_t *subsequences, const uint64_t *blocks, uint64_t n,
                      uint32_t *words)
{
    return blocks_hook(e, seed, subsequences, blocks, n, words, true);
}

int mcb_boxmuller_scan(mcb_engine *e, int which, uint64_t first_word, uint64_t count, double *max_abs_error,
                       uint64_t *n_bad)
{
    if (!e || !max_abs_error || !n_bad) return fail(MCB_ERR_INVALID, "NULL argument");
    if (which < 0 || which > 2) return fail(MCB_ERR_INVALID, "which must be 0 (radius), 1 (sin) or 2 (cos)");
    DeviceGuard g(e->device);
    int rc;
    if ((rc = e->scratch.reserve(16))) return rc;
    unsigned long long *d = reinterpret_cast<unsigned long long *>(e->scratch.ptr);
    CU(cudaMemsetAsync(d, 0, 16, e->stream));
    if (count) {
        boxmuller_scan_kernel<<<e->prop.multiProcessorCount * 8, 256, 0, e->stream>>>(which, first_word, count, d);
        e->launches++;
        CU(cudaGetLastError());
    }
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memcpy(max_abs_error, &h[0], sizeof(double));
    *n_bad = h[1];
    return MCB_OK;
}

int mcb_stream_normals(mcb_engine *e, uint64_t seed, uint64_t subsequence, uint64_t n0, uint64_t count, float *normals)
{
    if (!e || !normals) return fail(MCB_ERR_INVALID, "NULL argument");
    if (count == 0) return MCB_OK;
    DeviceGuard g(e->device);
    int rc;
    if ((rc = e->scratch.reserve((size_t)count * sizeof(float)))) return rc;
    float *d = reinterpret_cast<float *>(e->scratch.ptr);
    stream_normals_kernel<<<(unsigned)((count + 127) / 128), 128, 0, e->stream>>>(make_philox_keys(seed), subsequence,
                                                                               n0, count, d);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(normals, d, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_european_payoffs(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths,
                         uint64_t seed, int option_type, float *payoffs)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (!payoffs) return fail(MCB_ERR_INVALID, "payoffs is NULL");
    if (n_paths == 0) return MCB_OK;
    DeviceGuard g(e->device);
    const uint64_t c_lo = first_path / kEuropeanChunk;
    const uint64_t c_hi = (first_path + n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    if ((rc = e->scratch.reserve((size_t)n_paths * sizeof(float)))) return rc;
    float *d = reinterpret_cast<float *>(e->scratch.ptr);
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, first_path + n_paths, seed, c_lo);
    if ((rc = launch_european<MCB_EUROPEAN_PATHS_PER_SLOT>(e, prm, option_type, c_hi - c_lo, e->partials.ptr, d,
                                                           first_path, e->stream)))
        return rc;
    CU(cudaMemcpyAsync(payoffs, d, (size_t)n_paths * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_european_chunk_partials(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                                int option_type, float *partials, uint64_t n_chunks)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (!partials) return fail(MCB_ERR_INVALID, "partials is NULL");
    n_paths = resolve_paths(opt, n_paths);
    const uint64_t need = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    if (n_chunks != need) return fail(MCB_ERR_INVALID, "n_chunks must be %llu", (unsigned long long)need);
    DeviceGuard g(e->device);
    if ((rc = e->partials.reserve((size_t)need + 1))) return rc;
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, 0);
    if ((rc = launch_european<MCB_EUROPEAN_PATHS_PER_SLOT>(e, prm, option_type, need, e->partials.ptr, nullptr, 0,
                                                           e->stream)))
        return rc;
    CU(cudaMemcpyAsync(partials, e->partials.ptr, (size_t)need * sizeof(float2), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_bullet_payoffs(mcb_engine *e, const mcb_option_data *opt, uint64_t first_path, uint64_t n_paths, uint64_t seed,
                       int Ik, float Sk, int Tk, float *payoffs)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (!payoffs) return fail(MCB_ERR_INVALID, "payoffs is NULL");
    if (n_paths == 0) return MCB_OK;
    DeviceGuard g(e->device);
    const uint64_t c_lo = first_path / kBulletChunk;
    const uint64_t c_hi = (first_path + n_paths + kBulletChunk - 1) / kBulletChunk;
    if (c_hi - c_lo > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    if ((rc = e->scratch.reserve((size_t)n_paths * sizeof(float)))) return rc;
    float *d = reinterpret_cast<float *>(e->scratch.ptr);
    WalkParams prm;
    if ((rc = bullet_params(opt, first_path + n_paths, seed, Ik, Sk, Tk, c_lo, &prm))) return rc;
    bullet_kernel<MCB_BULLET_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, walk_table_bytes(prm), e->stream>>>(
        prm, e->partials.ptr, d, first_path);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(payoffs, d, (size_t)n_paths * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_last_segments(mcb_engine *e, double *segments)
{
    if (!e || !segments) return fail(MCB_ERR_INVALID, "NULL argument");
    if (e->last_job_chunks == 0) {   // bullet / sweep: copied home with the result
        memcpy(segments, e->h_segments, sizeof(double) * 2 * MCB_SEGMENTS);
        return MCB_OK;
    }
    // European job: its segments are still in the mailbox slot of its epoch (until kRing jobs later)
    if (e->job_epoch - e->last_job_epoch >= (unsigned long long)kRing)
        return fail(MCB_ERR_INVALID, "the segments of the last collected job have been overwritten");
    DeviceGuard g(e->device);
    const double *src = e->last_job_own ? e->mailbox->own[e->last_job_epoch % kRing] : e->mailbox->gather[e->last_job_epoch % kRing];
    CU(cudaMemcpy(segments, src, sizeof(double) * 2 * MCB_SEGMENTS, cudaMemcpyDeviceToHost));
    for (int sg = 0; sg < MCB_SEGMENTS; ++sg)   // segments without a chunk are +0.0 by rule (never stored)
        if ((e->last_job_chunks * (uint64_t)sg) / MCB_SEGMENTS == (e->last_job_chunks * (uint64_t)(sg + 1)) / MCB_SEGMENTS)
            segments[2 * sg] = segments[2 * sg + 1] = 0.0;
    return MCB_OK;
}

}  // extern "C"
