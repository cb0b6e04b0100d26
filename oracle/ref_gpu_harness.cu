// ref_gpu_harness.cu -- ORACLE / TEST INFRASTRUCTURE ONLY.
// A main() around the UNMODIFIED reference headers (compiled where they lie, -I/root/reference/inc,
// see oracle/Makefile) that runs the reference's OWN GPU wrappers on the B200 as a statistical
// live oracle:
//   wrapper_gpu_option_vanilla   inc/wrappers.cuh:33-57   (XORWOW, float atomics)
//   wrapper_gpu_bullet_option    inc/wrappers.cuh:59-93   (per-block partials, host sum)
// N_PATHS is forced to a multiple of threadsPerBlock (1024): with every thread in range the
// reference's known defects on this path (states allocated for N_PATHS but initialised for
// ceil(N/tpb)*tpb threads, barriers inside `if (idx < N_PATHS)`) cannot trigger.
// Its vanilla accumulator is never zeroed (inc/wrappers.cuh:43-47); a fresh process gets zeroed
// device pages, and the caller treats an absurd price as "reference defect", not as a mismatch.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "monte_carlo.cuh"

int main(int argc, char **argv)
{
    OptionData od;
    od.S0 = 100.0f; od.T = 1.0f; od.K = 100.0f; od.r = argc > 2 ? (float)atof(argv[2]) : 0.1f; od.v = 0.2f;
    od.B = 120.0f; od.P1 = 10; od.P2 = 50;
    od.N_PATHS = (argc > 1 ? atoi(argv[1]) : (1 << 20)) / 1024 * 1024;
    od.N_PATHS_INNER = 1; od.N_STEPS = 100;
    od.step = od.T / static_cast<float>(od.N_STEPS);
    cudaMemcpyToSymbol(d_OptionData, &od, sizeof(OptionData));   // as hello.cu:22
    cudaFree(0);   // context creation is not the wrapper's cost
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    const float vanilla = wrapper_gpu_option_vanilla(od, 1024);
    const auto t1 = clk::now();
    const float bullet = wrapper_gpu_bullet_option(od, 1024);
    const auto t2 = clk::now();
    printf("REFGPU %d %.9g %.9g\n", od.N_PATHS, vanilla, bullet);
    // whole-wrapper wall time, the way a caller of the reference experiences it: cudaMalloc of the
    // XORWOW states, setup_kernel, the pricing kernel, sync, D2H, cudaFree (inc/wrappers.cuh:36-56)
    printf("REFGPU_TIME %d %.6f %.6f\n", od.N_PATHS, std::chrono::duration<double>(t1 - t0).count(),
           std::chrono::duration<double>(t2 - t1).count());
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
