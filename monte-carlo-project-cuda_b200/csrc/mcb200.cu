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
#include <new>
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

struct mcb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    DeviceBuffer<float2> partials;
    DeviceBuffer<double> segments;
    DeviceBuffer<ResultDev> results;
    DeviceBuffer<unsigned char> scratch;   // hooks / host<->device staging
    DeviceBuffer<float> nested_ws;         // nested MC: log2 S, (prices), (counts) of the outer points
    DeviceBuffer<float> traj_ws;           // mcb_simulate_trajectories to a host buffer: one slab of rows (+ counts)
    DeviceBuffer<float4> sweep_sets;       // sweep: (c0, c1, K, -) per parameter set
    std::vector<float4> h_sweep_sets;      // host staging for sweep_sets (pageable on purpose)
    // peer-memory exchange (mcb_peer_mailbox_*): my mailbox, the peers' mailboxes mapped over CUDA IPC
    PeerMailbox *mailbox = nullptr;
    PeerTable peers{};
    void *peer_mapped[kMaxPeers] = {};     // what cudaIpcOpenMemHandle returned (to close on destroy)
    int peer_rank = -1, peer_world = 0;
    unsigned long long peer_epoch = 0;
    DeviceBuffer<unsigned int> seg_tickets; // fused peer kernel: per-segment arrival counters (zero between launches)
    mcb_result *h_results = nullptr;       // pinned
    size_t h_results_cap = 0;
    double *h_segments = nullptr;          // pinned, [MCB_SEGMENTS][2] of the last whole-job call
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
};

WalkConsts walk_consts(const mcb_option_data *o, double start)
{
    WalkConsts w;
    const double r = o->r, sig = o->v, dt = o->step;
    w.l0 = (float)std::log2(start);
    w.sc = (float)(sig * std::sqrt(dt) * kLog2e * kSqrt2Ln2d);
    w.dr = (float)((r - 0.5 * sig * sig) * dt * kLog2e);
    w.v = (float)(sig * std::sqrt(dt) * kLog2e);
    w.lB = o->B > 0.0f ? (float)std::log2((double)o->B) : -INFINITY;
    return w;
}

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
                    int seg_lo, int seg_hi, int n_sets, double *d_segments, cudaStream_t st)
{
    segment_kernel<<<dim3(MCB_SEGMENTS, (unsigned)n_sets), kSlots, 0, st>>>(partials, stride, first_chunk, n_chunks,
                                                                           seg_lo, seg_hi, d_segments);
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

// One row layout (SPL steps per lane, LPR lanes per row): the TMA slab kernel when the row fits one
// pass and is 16-byte aligned (the bandwidth path, config 3; counts and log2 prices ride along in
// their own staging rows), the general kernel otherwise.
template <int SPL, int LPR>
void launch_trajectory(const PathParams &prm, uint64_t n_paths, bool vec, bool base_aligned, float *d_prices,
                       int *d_counts, float *d_logs, cudaStream_t st)
{
    constexpr int kRowsPerWarp = 32 / LPR;
    constexpr int kSlabWarps = 4;
    // rows longer than one pass: one row group per slab, the whole rows staged -- while the staging
    // leaves enough CTAs per SM (measured on 2^20 rows: prices only 4.07 TB/s at 1024 steps / 32 KB,
    // 3.67 at 1536 / 48 KB, then the general kernel's 3.5 TB/s wins; with counts the general
    // kernel's direct stores are slow enough that staging pays up to 72 KB); longer or unaligned
    // rows take the general kernel
    const int n_arrays = 1 + (d_counts ? 1 : 0) + (d_logs ? 1 : 0);
    const size_t multi_smem = (size_t)kSlabWarps * n_arrays * kRowsPerWarp * (size_t)prm.n_steps * sizeof(float);
#define MCB_SLAB(ROWS, CNT, LOG, MULTI, ALIGNED)                                                              \
    do {                                                                                                      \
        auto kern = trajectory_slab_kernel<SPL, LPR, ROWS, kSlabWarps, CNT, LOG, MULTI, ALIGNED>;             \
        const uint64_t rows_per_cta = (uint64_t)kSlabWarps * ROWS;                                            \
        const uint64_t ctas = (n_paths + rows_per_cta - 1) / rows_per_cta;                                    \
        const size_t smem = (size_t)kSlabWarps * (1 + (CNT ? 1 : 0) + (LOG ? 1 : 0)) * ROWS *                 \
                            (size_t)prm.n_steps * sizeof(float);                                              \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<(unsigned)ctas, kSlabWarps * 32, smem, st>>>(prm, d_prices, d_counts, d_logs);                 \
    } while (0)
    // (only the largest layout is ever asked for rows longer than its pass)
    if (SPL * LPR == 256 && vec && prm.n_steps > SPL * LPR && multi_smem <= (n_arrays == 1 ? 48 : 72) * 1024) {
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
        if (d_counts && d_logs) MCB_SLAB(kRows3, true, true, false, true);
        else if (d_counts) MCB_SLAB(kRows2, true, false, false, true);
        else if (d_logs) MCB_SLAB(kRows2, false, true, false, true);
        else MCB_SLAB(kRows1, false, false, false, true);
    } else if (base_aligned && prm.n_steps <= SPL * LPR) {
        // rows that are not a multiple of 4 floats (150, 250 steps ...): slabs of 8 / 4 rows start on
        // 16-byte boundaries, so the bulk store still applies; only the staging is element-wise
        if (d_counts && d_logs) MCB_SLAB(4, true, true, false, false);
        else if (d_counts) MCB_SLAB(4, true, false, false, false);
        else if (d_logs) MCB_SLAB(4, false, true, false, false);
        else MCB_SLAB(8, false, false, false, false);
#undef MCB_SLAB
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

int mcb_engine_create(int device, mcb_engine **out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
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
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMallocHost(&e->h_segments, sizeof(double) * (2 * MCB_SEGMENTS + 8));
    if (err != cudaSuccess) {
        int rc = fail(MCB_ERR_CUDA, "engine setup failed: %s", cudaGetErrorString(err));
        mcb_engine_destroy(e);
        return rc;
    }
    memset(e->h_segments, 0, sizeof(double) * (2 * MCB_SEGMENTS + 8));
    if (e->segments.reserve(2 * MCB_SEGMENTS + 8) || reserve_results(e, 1)) {
        mcb_engine_destroy(e);
        return MCB_ERR_NOMEM;
    }
    *out = e;
    return MCB_OK;
}

int mcb_engine_destroy(mcb_engine *e)
{
    if (!e) return MCB_OK;
    DeviceGuard g(e->device);
    if (e->stream) {
        cudaStreamSynchronize(e->stream);
        cudaStreamDestroy(e->stream);
    }
    for (auto *v : {&e->timed, &e->event_pool})
        for (auto &t : *v) {
            cudaEventDestroy(t.start);
            cudaEventDestroy(t.stop);
        }
    e->partials.release();
    e->segments.release();
    e->results.release();
    e->scratch.release();
    e->nested_ws.release();
    e->traj_ws.release();
    e->sweep_sets.release();
    for (int r = 0; r < kMaxPeers; ++r)
        if (e->peer_mapped[r]) cudaIpcCloseMemHandle(e->peer_mapped[r]);
    if (e->mailbox) cudaFree(e->mailbox);
    e->seg_tickets.release();
    if (e->h_results) cudaFreeHost(e->h_results);
    if (e->h_segments) cudaFreeHost(e->h_segments);
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

int mcb_synchronize(mcb_engine *e)
{
    if (!e) return fail(MCB_ERR_INVALID, "engine is NULL");
    DeviceGuard g(e->device);
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

// --------------------------------------------------------------- peer-memory exchange (NVLink)
static_assert(MCB_MAX_PEERS == kMaxPeers && MCB_IPC_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "peer ABI");

int mcb_peer_mailbox_create(mcb_engine *e, void *handle_out)
{
    if (!e || !handle_out) return fail(MCB_ERR_INVALID, "NULL argument");
    DeviceGuard g(e->device);
    if (!e->mailbox) {
        CU(cudaMalloc(&e->mailbox, sizeof(PeerMailbox)));
        CU(cudaMemset(e->mailbox, 0, sizeof(PeerMailbox)));
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->mailbox));
    memcpy(handle_out, &h, sizeof(h));
    return MCB_OK;
}

int mcb_peer_mailbox_connect(mcb_engine *e, int rank, int world, const void *all_handles)
{
    if (!e || !all_handles) return fail(MCB_ERR_INVALID, "NULL argument");
    if (!e->mailbox) return fail(MCB_ERR_INVALID, "call mcb_peer_mailbox_create first");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(MCB_ERR_INVALID, "bad rank/world %d/%d (at most %d peers)", rank, world, kMaxPeers);
    if (MCB_SEGMENTS % world != 0) return fail(MCB_ERR_INVALID, "world must divide %d", MCB_SEGMENTS);
    DeviceGuard g(e->device);
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            e->peers.box[r] = e->mailbox;
            continue;
        }
        if (e->peer_mapped[r]) {
            cudaIpcCloseMemHandle(e->peer_mapped[r]);
            e->peer_mapped[r] = nullptr;
        }
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess));
        e->peer_mapped[r] = p;
        e->peers.box[r] = static_cast<PeerMailbox *>(p);
    }
    e->peer_rank = rank;
    e->peer_world = world;
    // peer_epoch is NOT reset: the mailbox (and the epochs already published in it) outlives a
    // re-connect, and the flags are compared with "<", so epochs must stay monotonic per engine
    return MCB_OK;
}

int mcb_european_peer_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                            int option_type, mcb_result *d_results, void *stream)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (e->peer_world < 1) return fail(MCB_ERR_INVALID, "peer mailboxes are not connected");
    if (!d_results) return fail(MCB_ERR_INVALID, "d_results is NULL");
    if (option_type != MCB_CALL && option_type != MCB_PUT) return fail(MCB_ERR_INVALID, "bad option_type");
    n_paths = resolve_paths(opt, n_paths);
    if (n_paths == 0) return fail(MCB_ERR_INVALID, "n_paths must be > 0");
    DeviceGuard g(e->device);
    cudaStream_t st = pick(e, stream);
    const int rank = e->peer_rank, world = e->peer_world;
    const uint64_t n_chunks = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
    if ((rc = launch_european<MCB_EUROPEAN_PATHS_PER_SLOT>(e, prm, option_type, c_hi - c_lo, e->partials.ptr, nullptr,
                                                           0, st)))
        return rc;
    const unsigned long long epoch = ++e->peer_epoch;
    segment_peer_kernel<<<(unsigned)(seg_hi - seg_lo), kSlots, 0, st>>>(e->partials.ptr, c_lo, n_chunks, seg_lo, seg_hi,
                                                                        e->peers, rank, world, epoch);
    e->launches++;
    CU(cudaGetLastError());
    const double discount = std::exp(-(double)opt->r * (double)opt->T);
    combine_peer_kernel<<<1, 32, 0, st>>>(e->mailbox, world, epoch, n_paths, discount,
                                          reinterpret_cast<ResultDev *>(d_results));
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

int mcb_european_fused_peer_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed,
                                  int option_type, mcb_result *d_results, void *stream)
{
    int rc = check_common(e, opt);
    if (rc) return rc;
    if (e->peer_world < 1) return fail(MCB_ERR_INVALID, "peer mailboxes are not connected");
    if (!d_results) return fail(MCB_ERR_INVALID, "d_results is NULL");
    if (option_type != MCB_CALL && option_type != MCB_PUT) return fail(MCB_ERR_INVALID, "bad option_type");
    n_paths = resolve_paths(opt, n_paths);
    const uint64_t n_chunks = (n_paths + kEuropeanChunk - 1) / kEuropeanChunk;
    // every segment needs at least one chunk for its "last CTA" to exist: small jobs take the 3-launch path
    if (n_chunks < (uint64_t)MCB_SEGMENTS)
        return mcb_european_peer_async(e, opt, n_paths, seed, option_type, d_results, stream);
    DeviceGuard g(e->device);
    cudaStream_t st = pick(e, stream);
    const int rank = e->peer_rank, world = e->peer_world;
    int seg_lo, seg_hi;
    uint64_t c_lo, c_hi;
    segment_span(rank, world, n_chunks, &seg_lo, &seg_hi, &c_lo, &c_hi);
    if (c_hi - c_lo > 0x7fffffffull) return fail(MCB_ERR_INVALID, "too many chunks for one launch");
    if ((rc = e->partials.reserve((size_t)(c_hi - c_lo) + 1))) return rc;
    if (e->seg_tickets.cap == 0) {
        if ((rc = e->seg_tickets.reserve(MCB_SEGMENTS))) return rc;
        CU(cudaMemsetAsync(e->seg_tickets.ptr, 0, sizeof(unsigned int) * MCB_SEGMENTS, st));
    }
    const EuropeanParams prm = european_params(opt, opt->K, opt->v, n_paths, seed, c_lo);
    FusedPeerArgs args{};
    args.n_chunks = n_chunks;
    args.n_paths = n_paths;
    args.discount = std::exp(-(double)opt->r * (double)opt->T);
    args.seg_tickets = e->seg_tickets.ptr;
    args.peers = e->peers;
    args.out = reinterpret_cast<ResultDev *>(d_results);
    args.epoch = ++e->peer_epoch;
    args.seg_lo = seg_lo; args.seg_hi = seg_hi; args.rank = rank; args.world = world;
    {
        TimedScope timed(e, MCB_KERNEL_EUROPEAN, st);
        if (option_type == MCB_PUT)
            european_fused_peer_kernel<kPut, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(
                prm, args, e->partials.ptr);
        else
            european_fused_peer_kernel<kCall, MCB_EUROPEAN_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(
                prm, args, e->partials.ptr);
    }
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

uint64_t mcb_launch_count(mcb_engine *e) { return e ? e->launches : 0; }

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
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    int rc = check_common(e, opt);
    if (rc) return rc;
    DeviceGuard g(e->device);
    n_paths = resolve_paths(opt, n_paths);
    if ((rc = mcb_european_segments_async(e, opt, n_paths, seed, option_type, 0, 1, e->segments.ptr, nullptr)))
        return rc;
    return finish_whole_job(e, 1, n_paths, opt->r, opt->T, out);
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
    p.l0 = w.l0; p.sc = w.sc; p.dr = w.dr; p.lB = w.lB;
    p.K = opt->K; p.P1 = opt->P1; p.P2 = opt->P2;
    p.n_steps = opt->N_STEPS - Tk;
    p.count0 = Ik;
    p.n_paths = n_paths_end;
    p.first_chunk = first_chunk;
    p.keys = make_philox_keys(seed);
    *out = p;
    return MCB_OK;
}

int mcb_bullet_segments_async(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int Ik,
                              float Sk, int Tk, int rank, int world, double *d_segments, void *stream)
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
            bullet_kernel<MCB_BULLET_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, st>>>(prm, e->partials.ptr,
                                                                                                 nullptr, 0);
        }
        e->launches++;
        CU(cudaGetLastError());
    }
    return launch_segments(e, e->partials.ptr, 0, c_lo, n_chunks, seg_lo, seg_hi, 1, d_segments, st);
}

int mcb_price_bullet(mcb_engine *e, const mcb_option_data *opt, uint64_t n_paths, uint64_t seed, int Ik, float Sk,
                     int Tk, mcb_result *out)
{
    if (!out) return fail(MCB_ERR_INVALID, "out is NULL");
    int rc = check_common(e, opt);
    if (rc) return rc;
    DeviceGuard g(e->device);
    n_paths = resolve_paths(opt, n_paths);
    if ((rc = mcb_bullet_segments_async(e, opt, n_paths, seed, Ik, Sk, Tk, 0, 1, e->segments.ptr, nullptr))) return rc;
    return finish_whole_job(e, 1, n_paths, opt->r, opt->T, out);
}

// ------------------------------------------------------------------------------------ sweep
int mcb_sweep_segments_async(mcb_engine *e, const mcb_option_data *opt, const float *strikes, const float *vols,
                             int n_params, uint64_t n_paths, uint64_t seed, int option_type, int rank, int world,
                             double *d_segments, void *stream)
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
        // A rank that owns fewer than ~3 waves of chunks (2 CTAs resident per SM at 128 registers) loses
        // its last, partly filled wave: halving the sets per CTA over gridDim.y doubles the CTA count
        // (measured at 512 chunks x 1024 sets: 2.71 -> 2.37 ms; finer splits and larger grids gain nothing)
        const uint64_t wave = 2ull * (uint64_t)e->prop.multiProcessorCount;
        uint64_t splits = (local && local < 3 * wave) ? 2 : 1;
        if (splits > (cnt + 63) / 64) splits = (cnt + 63) / 64;
        if (splits < 1) splits = 1;
        prm.sets_per_cta = (int)(((cnt + splits - 1) / splits + kSweepTile - 1) / kSweepTile * kSweepTile);
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
                                  d_segments + i0 * 2 * MCB_SEGMENTS, st)))
            return rc;
    }
    return MCB_OK;
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
    if ((rc = mcb_sweep_segments_async(e, opt, strikes, vols, n_params, n_paths, seed, option_type, 0, 1,
                                       e->segments.ptr, nullptr)))
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
    NestedParams prm{};
    prm.sc = w.sc; prm.dr = w.dr; prm.lB = w.lB;
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
        nested_kernel<<<dim3((unsigned)n_outer, (unsigned)split), kSlots, 0, st>>>(prm, ws_logs, ws_counts, d_F);
    }
    e->launches++;
    CU(cudaGetLastError());
    return MCB_OK;
}

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
    bullet_kernel<MCB_BULLET_PATHS_PER_SLOT><<<(unsigned)(c_hi - c_lo), kSlots, 0, e->stream>>>(prm, e->partials.ptr, d,
                                                                                                first_path);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(payoffs, d, (size_t)n_paths * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MCB_OK;
}

int mcb_last_segments(mcb_engine *e, double *segments)
{
    if (!e || !segments) return fail(MCB_ERR_INVALID, "NULL argument");
    memcpy(segments, e->h_segments, sizeof(double) * 2 * MCB_SEGMENTS);
    return MCB_OK;
}

}  // extern "C"
