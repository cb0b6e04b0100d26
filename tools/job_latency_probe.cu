// job_latency_probe.cu -- where the time of a small synchronous mcb_price_european call goes.  The whole engine is
// compiled into this probe with MCB_TRACE defined, so thread 0 of CTA 0 stamps clock64() at fixed points of the
// one-launch job (the stamps cost a few cycles each and are absent from libmcb200.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/job_latency_probe tools/job_latency_probe.cu
#include <cuda_runtime.h>
__device__ long long g_trace[16];
#define MCB_TRACE(i) if (threadIdx.x == 0 && blockIdx.x == 0) g_trace[i] = clock64();
__device__ unsigned long long g_cta[4][512];   // [stamp][CTA]: entry, payoffs sent, first barrier passed, slot sums done
#define MCB_TRACE_CTA(i)                                                              \
    if (threadIdx.x == 0) {                                                           \
        unsigned long long now;                                                       \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));                       \
        g_cta[i][blockIdx.x] = now;                                                   \
    }
#include "../monte-carlo-project-cuda_b200/csrc/mcb200.cu"

int main(int argc, char **argv)
{
    mcb_engine *e = nullptr;
    if (mcb_engine_create(0, &e)) { printf("%s\n", mcb_last_error()); return 1; }
    mcb_option_data opt{};
    opt.S0 = 100.f; opt.K = 100.f; opt.T = 1.f; opt.r = 0.1f; opt.v = 0.2f; opt.B = 0.f; opt.P1 = 0; opt.P2 = 0; opt.N_PATHS = 0; opt.N_STEPS = 1;
    unsigned long long sizes[16] = {1ull, 16384ull, 100000ull, 1000000ull, 2000000ull};   // or the sizes on the command line
    int n_sizes = 5;
    if (argc > 1) {
        n_sizes = argc - 1 > 16 ? 16 : argc - 1;
        for (int i = 0; i < n_sizes; ++i) sizes[i] = strtoull(argv[i + 1], nullptr, 10);
    }
    // small jobs (european_small_job_kernel): 1 = pricing + first cluster barrier, 2 = slot sums + fold + second barrier
    const char *names[10] = {"entry", "pricing", "fold", "segment ticket", "shard ticket", "tree loads+fold", "fp64 stats",
                             "result stores", "", ""};
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    for (int k = 0; k < n_sizes; ++k) {
        mcb_result r;
        const int reps = 3000;
        for (int i = 0; i < 300; ++i) mcb_price_european(e, &opt, sizes[k], 1234, MCB_CALL, &r);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) mcb_price_european(e, &opt, sizes[k], 1234, MCB_CALL, &r);
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        cudaDeviceSynchronize();
        long long t[16];
        cudaMemcpyFromSymbol(t, g_trace, sizeof(t));
        printf("%llu paths: %.2f us per call (price %.6f); CTA 0 inside the kernel, cycles since entry (max clock %d MHz):\n",
               sizes[k], us, r.price, clk_khz / 1000);
        long long prev = t[0];
        for (int i = 1; i <= 7; ++i)
            if (t[i] > t[0]) {
                printf("   after %-16s %7lld  (+%lld)\n", names[i], t[i] - t[0], t[i] - prev);
                prev = t[i];
            }
        {
            long long zero[16] = {0};
            cudaMemcpyToSymbol(g_trace, zero, sizeof(zero));
        }
        if (sizes[k] <= 1000000ull) {
            static unsigned long long c[4][512];
            cudaMemcpyFromSymbol(c, g_cta, sizeof(c));
            const int n_cta = (int)((sizes[k] + 16383) / 16384) * 8;
            unsigned long long first = ~0ull;
            for (int i = 0; i < n_cta; ++i) first = c[0][i] < first ? c[0][i] : first;
            const char *what[4] = {"entry", "payoffs sent", "first barrier passed", "slot sums done"};
            for (int st = 0; st < 4; ++st) {
                unsigned long long lo = ~0ull, hi = 0, sum = 0;
                for (int i = 0; i < n_cta; ++i) {
                    const unsigned long long v = c[st][i] - first;
                    lo = v < lo ? v : lo; hi = v > hi ? v : hi; sum += v;
                }
                printf("   all %d CTAs, ns after the first entry: %-22s min %5llu  mean %5llu  max %5llu\n", n_cta, what[st], lo,
                       sum / n_cta, hi);
            }
        }
        // host side alone: submit without waiting, then drain
        uint64_t tk[4];
        const auto h0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) {
            mcb_european_submit(e, &opt, sizes[k], 1234, MCB_CALL, &tk[i & 3]);
            if ((i & 3) == 3) for (int j = 0; j < 4; ++j) mcb_european_collect(e, tk[j], &r);
        }
        printf("   pipelined (4 in flight): %.2f us per call\n",
               std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count() / reps);
    }
    mcb_engine_destroy(e);
    return 0;
}
