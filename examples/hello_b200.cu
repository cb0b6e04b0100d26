// hello_b200.cu -- the reference's canonical caller sequence (hello.cu:3-47: fill OptionData, upload
// d_OptionData, print, call the eight wrappers, print the closed form) written against the compat
// headers.  The reference's own hello.cu compiles unchanged the same way:
//   nvcc -std=c++17 -Iinclude/compat examples/hello_b200.cu -Lmonte-carlo-project-cuda_b200 -lmcb200
#include "monte_carlo.cuh"

int main(int argc, char **argv)
{
    OptionData od{};
    od.S0 = 100.0f; od.T = 1.0f; od.K = 100.0f; od.r = 0.1f; od.v = 0.2f;
    od.B = 120.0f; od.P1 = 10; od.P2 = 50;
    od.N_PATHS = argc > 1 ? atoi(argv[1]) : 100000;
    od.N_PATHS_INNER = argc > 2 ? atoi(argv[2]) : 1000;
    od.N_STEPS = 100;
    od.step = od.T / (float)od.N_STEPS;
    const int tpb = 1024;

    cudaMemcpyToSymbol(d_OptionData, &od, sizeof(OptionData));  // accepted and ignored by the new engine
    printOptionData(od);
    getDeviceProperty();

    const float cpu_vanilla = wrapper_cpu_option_vanilla(od, tpb);
    const float cpu_bullet = wrapper_cpu_bullet_option(od, tpb);
    const float gpu_vanilla = wrapper_gpu_option_vanilla(od, tpb);
    const float gpu_bullet = wrapper_gpu_bullet_option(od, tpb);
    const float gpu_bullet_atomic = wrapper_gpu_bullet_option_atomic(od, tpb);

    OptionData nested = od;  // the nested wrappers cost N_PATHS * N_STEPS * N_PATHS_INNER * N_STEPS / 2
    nested.N_PATHS = od.N_PATHS < 2048 ? od.N_PATHS : 2048;
    const float n1 = wrapper_gpu_bullet_option_nmc_one_point_one_block(nested, tpb, 5000);
    const float n2 = wrapper_gpu_bullet_option_nmc_one_kernel(nested, tpb, 5000);
    const float n3 = wrapper_gpu_bullet_option_nmc_optimal(nested, tpb, 5000);

    float closed = 0.0f;
    black_scholes_CPU(closed, od.S0, od.K, od.T, od.r, od.v);
    cout << endl << "call Black Scholes : " << closed << endl;

    // machine-readable tail for tests/test_compat_headers.py
    printf("RESULT %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", cpu_vanilla, cpu_bullet, gpu_vanilla, gpu_bullet,
           gpu_bullet_atomic, n1, n2, n3, closed);
    return (gpu_vanilla < 0 || gpu_bullet < 0 || n1 < 0) ? 1 : 0;
}
