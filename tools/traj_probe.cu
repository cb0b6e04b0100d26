// traj_probe.cu -- where does the trajectory kernel's time go?  Times stripped-down variants of the
// per-pass work (same device functions as the product, csrc/philox.cuh) with stores replaced by a
// checksum.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I monte-carlo-project-cuda_b200/csrc ...
#include <cstdio>
#include <cuda_runtime.h>
#include "philox.cuh"
#include "block_reduce.cuh"
using namespace mcb;

enum { PHILOX = 1, BOXMULLER = 2, PREFIX = 4, SCAN = 8, EXP2 = 16, STORE = 32 };

template <int WHAT>
__global__ void __launch_bounds__(256) probe(PhiloxKeys keys, float sc, float dr, int passes, float *out, float *sink)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * 256 + threadIdx.x) >> 5;
    float acc = 0.0f;
    uint32_t xacc = 0;
    float carry = 6.64f;
    for (int r = 0; r < passes; ++r) {
        const uint32_t p_lo = gw * passes + r;
        float a[8];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            Words4 w;
            if (WHAT & PHILOX) w = philox4x32_10((uint32_t)(2 * lane + b), 0u, p_lo, 0u, keys);
            else { w.x = p_lo * 2654435761u + lane * 40503u + b; w.y = w.x * 2246822519u; w.z = w.y ^ (w.x >> 7); w.w = w.z * 3266489917u; }
            if (WHAT & BOXMULLER) increments4(w, sc, dr, a + 4 * b);
            else { a[4*b] = __uint_as_float((w.x >> 9) | 0x3c000000u); a[4*b+1] = __uint_as_float((w.y >> 9) | 0x3c000000u);
                   a[4*b+2] = __uint_as_float((w.z >> 9) | 0x3c000000u); a[4*b+3] = __uint_as_float((w.w >> 9) | 0x3c000000u); }
        }
        if (WHAT & PREFIX) {
#pragma unroll
            for (int j = 1; j < 8; ++j) a[j] = a[j] + a[j - 1];
        }
        float base = carry;
        if (WHAT & SCAN) {
            float x = __shfl_up_sync(kFullMask, a[7], 1);
            if (lane == 0) x = 0.0f;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { float y = __shfl_up_sync(kFullMask, x, off); if (lane >= off) x += y; }
            base += x;
        }
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float l = base + a[j]; s[j] = (WHAT & EXP2) ? mufu_ex2(l) : l; }
        if (WHAT & STORE) {
            float4 *dst = reinterpret_cast<float4 *>(out + ((size_t)p_lo * 256 + lane * 8));
            __stcs(dst, make_float4(s[0], s[1], s[2], s[3]));
            __stcs(dst + 1, make_float4(s[4], s[5], s[6], s[7]));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += s[j];
        }
    }
    if (acc == 123.456f || xacc == 77u) sink[0] = acc;
}

// Specialisation experiment: warps 0-3 of a CTA run Philox-only passes, warps 4-7 run everything else
// (Box-Muller, prefix, scan, exp2) on cheap integers -- no data exchange, just the two instruction
// streams side by side on every SMSP.  Same total number of passes of each kind as the fused kernel.
__global__ void __launch_bounds__(256) probe_split(PhiloxKeys keys, float sc, float dr, int passes, float *out, float *sink)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t pair = blockIdx.x * 4 + (warp & 3);
    float acc = 0.0f;
    uint32_t xacc = 0;
    if (warp < 4) {
        for (int r = 0; r < 2 * passes; ++r) {
            const uint32_t p_lo = pair * 2 * passes + r;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const Words4 w = philox4x32_10((uint32_t)(2 * lane + b), 0u, p_lo, 0u, keys);
                xacc ^= w.x ^ w.y ^ w.z ^ w.w;
            }
        }
    } else {
        float carry = 6.64f;
        for (int r = 0; r < 2 * passes; ++r) {
            const uint32_t p_lo = pair * 2 * passes + r;
            float a[8];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                Words4 w;
                w.x = p_lo * 2654435761u + lane * 40503u + b; w.y = w.x * 2246822519u; w.z = w.y ^ (w.x >> 7); w.w = w.z * 3266489917u;
                increments4(w, sc, dr, a + 4 * b);
            }
#pragma unroll
            for (int j = 1; j < 8; ++j) a[j] = a[j] + a[j - 1];
            float x = __shfl_up_sync(kFullMask, a[7], 1);
            if (lane == 0) x = 0.0f;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { float y = __shfl_up_sync(kFullMask, x, off); if (lane >= off) x += y; }
            const float base = carry + x;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += mufu_ex2(base + a[j]);
        }
    }
    if (acc == 123.456f || xacc == 77u) sink[0] = acc + xacc;
}

void run_split(float *out, float *sink)
{
    const int passes = 8, blocks = (1 << 20) / passes / 8;
    PhiloxKeys k = make_philox_keys(1234);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) probe_split<<<blocks, 256>>>(k, 0.0214f, 1.6e-4f, passes, out, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) probe_split<<<blocks, 256>>>(k, 0.0214f, 1.6e-4f, passes, out, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("%-44s %8.1f us  %6.3f Tsteps/s  %s\n", "SPLIT: int warps || fp warps (no exchange)", ms * 1e3,
           (double)(1 << 20) * 256 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

// Thread-per-path experiment: lane = row, serial walk (no prefix / scan), TS steps staged per tile in
// shared memory (padded rows, conflict-free STS.128), every lane hands its own row segment to the TMA
// engine (cp.async.bulk of TS*4 bytes), double-buffered tiles.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int TS>
__global__ void __launch_bounds__(128) probe_rows(PhiloxKeys keys, float sc, float dr, float *out)
{
    constexpr int kStride = TS + 4;
    __shared__ __align__(128) float tile[4][2][32][kStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row = (blockIdx.x * 4 + warp) * 32 + lane;
    float l = 6.64f;
    for (int t = 0; t < 256 / TS; ++t) {
        float *buf = &tile[warp][t & 1][lane][0];
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // my copy from two tiles ago has been read
#pragma unroll 2
        for (int b = 0; b < TS / 4; ++b) {
            float d[4];
            increments4(philox4x32_10((uint32_t)(t * (TS / 4) + b), 0u, row, 0u, keys), sc, dr, d);
            float4 s;
            l += d[0]; s.x = mufu_ex2(l);
            l += d[1]; s.y = mufu_ex2(l);
            l += d[2]; s.z = mufu_ex2(l);
            l += d[3]; s.w = mufu_ex2(l);
            *reinterpret_cast<float4 *>(buf + 4 * b) = s;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        float *dst = out + (size_t)row * 256 + t * TS;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(buf)), "r"(TS * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int TS>
void run_rows(float *out)
{
    const int blocks = (1 << 20) / 128;
    PhiloxKeys k = make_philox_keys(1234);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) probe_rows<TS><<<blocks, 128>>>(k, 0.0214f, 1.6e-4f, out);
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) probe_rows<TS><<<blocks, 128>>>(k, 0.0214f, 1.6e-4f, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("ROWS: thread per path, tile %3d steps + per-lane bulk %8.1f us  %6.3f Tsteps/s  %s\n", TS, ms * 1e3,
           (double)(1 << 20) * 256 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

template <int WHAT>
void run(const char *name, float *out, float *sink)
{
    const int passes = 8, blocks = (1 << 20) / passes / 8;  // 2^20 rows of 256 steps in total
    PhiloxKeys k = make_philox_keys(1234);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) probe<WHAT><<<blocks, 256>>>(k, 0.0214f, 1.6e-4f, passes, out, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) probe<WHAT><<<blocks, 256>>>(k, 0.0214f, 1.6e-4f, passes, out, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    const double steps = (double)(1 << 20) * 256;
    printf("%-44s %8.1f us  %6.3f Tsteps/s  %s\n", name, ms * 1e3, steps / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out, *sink;
    cudaMalloc(&out, (size_t)(1 << 20) * 256 * 4);
    cudaMalloc(&sink, 64);
    run<PHILOX>("philox only", out, sink);
    run<BOXMULLER>("box-muller only (cheap ints)", out, sink);
    run<PHILOX | BOXMULLER>("philox + box-muller", out, sink);
    run<PHILOX | BOXMULLER | PREFIX>("+ prefix", out, sink);
    run<PHILOX | BOXMULLER | PREFIX | SCAN>("+ scan", out, sink);
    run<PHILOX | BOXMULLER | PREFIX | SCAN | EXP2>("+ ex2 (all compute)", out, sink);
    run<PHILOX | BOXMULLER | PREFIX | SCAN | EXP2 | STORE>("+ stores (full)", out, sink);
    run<PHILOX | BOXMULLER | PREFIX | EXP2 | STORE>("full without scan", out, sink);
    run<BOXMULLER | PREFIX | SCAN | EXP2 | STORE>("full without philox", out, sink);
    run<BOXMULLER | PREFIX>("fp only: bm + prefix", out, sink);
    run<BOXMULLER | PREFIX | SCAN>("fp only: bm + prefix + scan", out, sink);
    run<BOXMULLER | PREFIX | EXP2>("fp only: bm + prefix + ex2 (no scan)", out, sink);
    run<BOXMULLER | PREFIX | SCAN | EXP2>("fp only: bm + prefix + scan + ex2", out, sink);
    run_split(out, sink);
    run_rows<16>(out);
    run_rows<32>(out);

    run<STORE>("stores only", out, sink);
    run<EXP2 | STORE>("ex2 + stores", out, sink);
    return 0;
}
