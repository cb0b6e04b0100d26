// launch_floor.cu -- the platform floor under a synchronous one-launch call: an EMPTY kernel whose only act is to
// write a sequence word into mapped pinned host memory, the host spinning on that word (exactly how
// mcb_european_collect waits).  mcb_price_european's 17-20 us per call is to be read against this number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_floor tools/launch_floor.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

struct Params { char pad[352]; };   // the job kernels carry ~350 bytes of parameters (option, Philox keys, peer table)

__global__ void flag_kernel(const __grid_constant__ Params p, volatile unsigned long long *flag, unsigned long long seq)
{
    if (threadIdx.x == 0 && p.pad[0] == 0) {
        __threadfence_system();
        *flag = seq;
    }
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
    return 0;
}
