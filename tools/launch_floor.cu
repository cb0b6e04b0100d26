// launch_floor.cu -- the platform floor under a synchronous one-launch call: an EMPTY kernel whose only act is to
// write a sequence word into mapped pinned host memory, the host spinning on that word (exactly how
// mcb_european_collect waits).  mcb_price_european's 17-20 us per call is to be read against this number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_floor tools/launch_floor.cu
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

struct Params { char pad[352]; };   // the job kernels carry ~350 bytes of parameters (option, Philox keys, peer table)

__global__ void flag_kernel(const __grid_constant__ Params p, volatile unsigned long long *flag, unsigned long long seq)
{
    if (threadIdx.x == 0 && p.pad[0] == 0) {
        __threadfence_system();
        *flag = seq;
    }
}

// ---- round 2, second study: what the tail of a one-launch pricing call adds to that floor -------------------------
struct Slot { double r[5]; unsigned long long seq; unsigned long long pad[2]; };
struct SlotLL { unsigned long long w[10]; };   // (data32, seq32) pairs: every 8-byte store is complete in itself

// (a) the shipped result protocol: five doubles, a system fence, the sequence word
__global__ void result_fenced_kernel(const __grid_constant__ Params p, volatile Slot *slot, unsigned long long seq)
{
    if (threadIdx.x == 0 && p.pad[0] == 0) {
        for (int i = 0; i < 5; ++i) slot->r[i] = (double)seq + i;
        __threadfence_system();
        slot->seq = seq;
    }
}
// (b) flag-in-data: ten lanes store (data word, seq) pairs, no fence
__global__ void result_ll_kernel(const __grid_constant__ Params p, volatile SlotLL *slot, unsigned long long seq)
{
    if (threadIdx.x < 10 && p.pad[0] == 0) {
        const double v = (double)seq + (threadIdx.x >> 1);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        const unsigned int half = (threadIdx.x & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
        slot->w[threadIdx.x] = ((unsigned long long)(unsigned int)seq << 32) | half;
    }
}
// (c) (a) + what the job tail does before it: two stores to a mailbox, fence, a ticket round trip, a barrier, 64
// volatile loads, the double-precision mean / variance / sqrt
__global__ void result_tail_kernel(const __grid_constant__ Params p, volatile Slot *slot, unsigned long long seq,
                                   double *mailbox, unsigned int *ticket, int with_math)
{
    __shared__ int last;
    if (p.pad[0] != 0) return;
    if (threadIdx.x == 0) {
        __stcg(mailbox, (double)seq);
        __stcg(mailbox + 1, 2.0 * (double)seq);
        __threadfence();
        last = atomicAdd(ticket, 1u) == 0u;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;
    if (threadIdx.x < 32) {
        volatile double *m = mailbox;
        double s = (threadIdx.x == 0 ? m[0] : 0.0), q = (threadIdx.x == 0 ? m[1] : 0.0);
        for (int o = 16; o; o >>= 1) {
            s += __shfl_down_sync(0xffffffffu, s, o);
            q += __shfl_down_sync(0xffffffffu, q, o);
        }
        if (threadIdx.x == 0) {
            double n = 1000.0 + (double)(seq & 7), mean = s, var = q, se = q;
            if (with_math) {
                mean = s / n;
                var = q / n - mean * mean;
                var = var > 0.0 ? var : 0.0;
                var *= n / (n - 1.0);
                se = sqrt(var / n);
            }
            slot->r[0] = mean; slot->r[1] = se; slot->r[2] = s; slot->r[3] = q; slot->r[4] = n;
            __threadfence_system();
            slot->seq = seq;
        }
    }
}

template <class Launch, class Done> static double spin_loop(int reps, unsigned long long base, Launch launch, Done done)
{
    double us = 0.0;
    for (int warm = 0; warm < 2; ++warm) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 1; i <= reps; ++i) {
            const unsigned long long seq = base + (unsigned long long)warm * reps + i;
            launch(seq);
            while (!done(seq)) {}
        }
        us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    }
    return us;
}

int main()
{
    unsigned long long *flag;
    cudaHostAlloc(&flag, 64, cudaHostAllocMapped);
    *flag = 0;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    Params p{};
    const int reps = 5000;
    for (int warm = 0; warm < 2; ++warm) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 1; i <= reps; ++i) {
            const unsigned long long seq = (unsigned long long)warm * reps + i;
            flag_kernel<<<1, 256, 0, st>>>(p, flag, seq);
            while (*(volatile unsigned long long *)flag != seq) {}
        }
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        if (warm) printf("empty kernel + host-visible flag: %.2f us per launch (host spin on mapped memory)\n", us);
    }
    // the same with cudaStreamSynchronize instead of the spin
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 1; i <= reps; ++i) {
        flag_kernel<<<1, 256, 0, st>>>(p, flag, 3ull * reps + i);
        cudaStreamSynchronize(st);
    }
    printf("empty kernel + cudaStreamSynchronize: %.2f us per launch\n",
           std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps);
    // ---- the tail study ----
    Slot *slot;
    SlotLL *ll;
    cudaHostAlloc(&slot, sizeof(Slot), cudaHostAllocMapped);
    cudaHostAlloc(&ll, sizeof(SlotLL), cudaHostAllocMapped);
    memset(slot, 0, sizeof(Slot));
    memset(ll, 0, sizeof(SlotLL));
    double *mailbox;
    unsigned int *ticket;
    cudaMalloc(&mailbox, 1024);
    cudaMalloc(&ticket, 4);
    cudaMemset(ticket, 0, 4);
    cudaDeviceSynchronize();
    const unsigned long long base = 100000ull;
    double us = spin_loop(reps, base, [&](unsigned long long seq) { result_fenced_kernel<<<1, 256, 0, st>>>(p, slot, seq); },
                          [&](unsigned long long seq) { return ((volatile Slot *)slot)->seq == seq; });
    printf("5 doubles + fence.sys + seq word:            %.2f us per launch\n", us);
    us = spin_loop(reps, base, [&](unsigned long long seq) { result_ll_kernel<<<1, 256, 0, st>>>(p, ll, seq); },
                   [&](unsigned long long seq) {
                       for (int i = 0; i < 10; ++i)
                           if ((unsigned int)(((volatile SlotLL *)ll)->w[i] >> 32) != (unsigned int)seq) return false;
                       return true;
                   });
    printf("10 x (data32, seq32) stores, no fence:        %.2f us per launch\n", us);
    for (int math = 0; math < 2; ++math) {
        us = spin_loop(reps, base * (2 + math), [&](unsigned long long seq) { result_tail_kernel<<<1, 256, 0, st>>>(p, slot, seq, mailbox, ticket, math); },
                       [&](unsigned long long seq) { return ((volatile Slot *)slot)->seq == seq; });
        printf("mailbox + ticket + tree%s + fenced result: %.2f us per launch\n", math ? " + fp64 stats" : "             ", us);
    }
    return 0;
}
