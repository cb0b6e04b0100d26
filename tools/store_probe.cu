// store_probe.cu -- pure-store bandwidth of the access patterns a trajectory kernel can use
// (rows of ROWF floats, one row per warp per pass, 2^20 rows).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

enum { COALESCED_CS, COALESCED_WB, LANE32B_CS, LANE32B_WB, BULK_1ROW, BULK_4ROWS, LANE64B_CS };

template <int MODE, int ROWF>
__global__ void __launch_bounds__(256) probe(float *out, int rows_per_warp)
{
    __shared__ __align__(128) float stage[8][2][(MODE == BULK_4ROWS ? 2 : 1) * 256 * (MODE >= BULK_1ROW && MODE <= BULK_4ROWS ? 1 : 0) + 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * 8 + warp;
    const float4 v = make_float4(lane, warp, gw, 1.0f);
    for (int r = 0; r < rows_per_warp; ++r) {
        const size_t row = (size_t)gw * rows_per_warp + r;
        float *base = out + row * ROWF;
        if (MODE == COALESCED_CS || MODE == COALESCED_WB) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int idx = k * 128 + lane * 4;
                if (idx < ROWF) {
                    if (MODE == COALESCED_CS) __stcs(reinterpret_cast<float4 *>(base + idx), v);
                    else *reinterpret_cast<float4 *>(base + idx) = v;
                }
            }
        } else if (MODE == LANE32B_CS || MODE == LANE32B_WB) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int idx = lane * 8 + 4 * k;
                if (idx < ROWF) {
                    if (MODE == LANE32B_CS) __stcs(reinterpret_cast<float4 *>(base + idx), v);
                    else *reinterpret_cast<float4 *>(base + idx) = v;
                }
            }
        } else if (MODE == LANE64B_CS) {   // 16 lanes per row, two rows per pass
            const int sub = lane >> 4, ln = lane & 15;
            float *b2 = out + ((size_t)gw * rows_per_warp + (r & ~1) + sub) * ROWF;
            if ((r & 1) == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int idx = ln * 16 + 4 * k;
                    if (idx < ROWF) __stcs(reinterpret_cast<float4 *>(b2 + idx), v);
                }
            }
        } else if (MODE == BULK_1ROW) {
            float *buf = stage[warp][r & 1];
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            *reinterpret_cast<float4 *>(buf + lane * 8) = v;
            *reinterpret_cast<float4 *>(buf + lane * 8 + 4) = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(base), "r"(smem_addr(buf)), "r"(ROWF * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (MODE == BULK_4ROWS) {   // stage 4 rows, one bulk copy of 4*ROWF*4 bytes (rows are contiguous)
            float *buf = stage[warp][(r >> 1) & 1];
            if ((r & 1) == 0) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            float *dst = buf + (r & 1) * ROWF + lane * 8;
            if (lane * 8 < ROWF) {
                *reinterpret_cast<float4 *>(dst) = v;
                if (lane * 8 + 4 < ROWF) *reinterpret_cast<float4 *>(dst + 4) = v;
            }
            if ((r & 1) == 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    float *g = out + ((size_t)gw * rows_per_warp + (r - 1)) * ROWF;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g), "r"(smem_addr(buf)), "r"(2 * ROWF * 4) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
    }
    if (MODE == BULK_1ROW || MODE == BULK_4ROWS) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

template <int MODE, int ROWF>
void run(const char *name, float *out, int rows_per_warp)
{
    const int rows = 1 << 20, blocks = rows / rows_per_warp / 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) probe<MODE, ROWF><<<blocks, 256>>>(out, rows_per_warp);
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) probe<MODE, ROWF><<<blocks, 256>>>(out, rows_per_warp);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
    printf("%-34s row=%d floats rows/warp=%d %8.1f us %8.1f GB/s  %s\n", name, ROWF, rows_per_warp, ms * 1e3,
           (double)rows * ROWF * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out;
    cudaMalloc(&out, (size_t)(1 << 20) * 256 * 4 + 4096);
    for (int rpw : {8, 32}) {
        run<COALESCED_CS, 256>("coalesced 512B/instr .cs", out, rpw);
        run<COALESCED_WB, 256>("coalesced 512B/instr default", out, rpw);
        run<LANE32B_CS, 256>("32B per lane (2 instr) .cs", out, rpw);
        run<LANE32B_WB, 256>("32B per lane (2 instr) default", out, rpw);
        run<LANE64B_CS, 256>("64B per lane (4 instr) .cs", out, rpw);
        run<BULK_1ROW, 256>("smem + bulk 1 row", out, rpw);
        run<BULK_4ROWS, 256>("smem + bulk 2 rows", out, rpw);
        run<COALESCED_CS, 252>("coalesced 512B/instr .cs", out, rpw);
        run<LANE32B_CS, 252>("32B per lane (2 instr) .cs", out, rpw);
        run<LANE32B_WB, 252>("32B per lane (2 instr) default", out, rpw);
        run<BULK_1ROW, 252>("smem + bulk 1 row", out, rpw);
        run<BULK_4ROWS, 252>("smem + bulk 2 rows", out, rpw);
    }
    return 0;
}
