// compat/testing.cuh -- drop-in for the reference's test-harness header (inc/testing.cuh) over
// libmcb200.so: enum ReductionType (:100-106), class Simulation (:108-405) with the same public
// members and methods, the pre-generated-normal CPU pricer (:75-91) and the random-array helpers
// (:17-42).  The reference's testing.cu compiles unchanged against this header.
// Differences a caller sees: the normals come from the engine's Philox stream (seed, subsequence
// 0) instead of cuRAND's XORWOW host generator; reductions and trajectories are deterministic.
#pragma once

#include <cmath>
#include <cstdint>
#include <iostream>
#include <vector>

#include "tool.cuh"
#include "reduce.cuh"
#include "trajectories.cuh"
#include "nmc.cuh"

// cuRAND-host-API replacement: `length` standard normals on the device and on the host.
inline void generate_random_array(float *d_randomData, float *h_randomData, int length, unsigned long long seed = 1234ULL)
{
    mcb_engine *e = mcb_compat_engine();
    if (!e || mcb_generate_normals(e, seed, (uint64_t)length, d_randomData, MCB_DEVICE) != MCB_OK) {
        fprintf(stderr, "mcb200: %s\n", mcb_last_error());
        exit(EXIT_FAILURE);   // as the reference's testCUDA would
    }
    testCUDA(cudaMemcpy(h_randomData, d_randomData, (size_t)length * sizeof(float), cudaMemcpyDeviceToHost));
}

inline void init_random_array(float **d_randomData, float **h_randomData, size_t length, long seed = 1234ULL)
{
    testCUDA(cudaMalloc(d_randomData, length * sizeof(float)));
    *h_randomData = (float *)malloc(length * sizeof(float));
    CHECK_MALLOC(*h_randomData);
    generate_random_array(*d_randomData, *h_randomData, (int)length, (unsigned long long)seed);
}

// CPU pricer from pre-generated normals h_randomData[path*N_STEPS + step] (inc/testing.cuh:75-91):
// per-path call payoffs into simulated_paths_cpu, their undiscounted mean into *optionPriceCPU.
inline void simulateOptionPriceCPU(float *optionPriceCPU, int N_PATHS, int N_STEPS, float *h_randomData, float S0,
                                   float sigma, float sqrdt, float r, float K, float dt, float *simulated_paths_cpu)
{
    const float drift = (r - 0.5f * sigma * sigma) * dt, vol = sigma * sqrdt;
    double total = 0.0;
    for (int path = 0; path < N_PATHS; ++path) {
        const float *z = h_randomData + (size_t)path * N_STEPS;
        float log_s = logf(S0);
        for (int k = 0; k < N_STEPS; ++k) log_s += drift + vol * z[k];
        const float terminal = expf(log_s);
        simulated_paths_cpu[path] = terminal > K ? terminal - K : 0.0f;
        total += simulated_paths_cpu[path];
    }
    *optionPriceCPU = (float)(total / N_PATHS);
}

enum ReductionType {
    SequentialAddressing = 3,
    FirstAddDuringLoad = 4,
    UnrollLastWarp = 5,
    CompletelyUnrolled = 6
};

class Simulation {
public:
    size_t n_trajectories;
    size_t n_steps;
    float *d_random_array = nullptr;
    float *h_random_array = nullptr;

    Simulation(size_t n_trajectories = 10, size_t n_steps = 100, float volatilty = 0.2, float risk_free_rate = 0.1,
               float initial_spot_price = 100.0, float contract_strike = 100.0, float contract_maturity = 1,
               float barrier = 0, float P1 = 0, float P2 = 0)
        : n_trajectories{n_trajectories}, n_steps{n_steps}, sigma{volatilty}, r{risk_free_rate},
          x_0{initial_spot_price}, K{contract_strike}, T{contract_maturity}, B{barrier}, P1{P1}, P2{P2}
    {
        this->initialize_random_array();
    }

    // The engine's generator is stateless: there is no per-thread RNG state to initialise.
    void initialize_rng_state(size_t /*threads_per_block*/, uint64_t /*seed*/ = 1234) {}

    float sum_random_array()
    {
        if (!h_random_array) initialize_random_array();
        float out = 0.0f;
        for (size_t i = 0; i < length(); ++i) out += h_random_array[i];
        return out;
    }

    // Per-block results with the index ranges of reduce3..6 at <<<n_blocks, n_threads_per_block>>>:
    // kinds 3-5 give block b the 2*threads elements starting at b*2*threads, kind 6 grid-strides.
    std::vector<float> test_reduction(size_t n_blocks, size_t n_threads_per_block, int reduction)
    {
        if (!d_random_array) initialize_random_array();
        std::vector<float> out(n_blocks, 0.0f);
        mcb_engine *e = mcb_compat_engine();
        const int strided = reduction == CompletelyUnrolled;
        if (!e || mcb_reduce_blocks(e, d_random_array, length(), MCB_DEVICE, (uint32_t)n_blocks,
                                    2 * (uint64_t)n_threads_per_block, strided, out.data()) != MCB_OK) {
            fprintf(stderr, "mcb200: %s\n", mcb_last_error());
            exit(EXIT_FAILURE);
        }
        return out;
    }

    std::vector<float> simulate_trajectory_cpu()
    {
        if (!d_random_array || !h_random_array) {
            std::cout << "Initializing random array with " << length() << " elements\n";
            initialize_random_array();
        }
        std::cout << "Simulating trajectories using reduction method: \n";
        float option_price = 0.0f;
        std::vector<float> payoffs(n_trajectories, 0.0f);
        simulateOptionPriceCPU(&option_price, (int)n_trajectories, (int)n_steps, h_random_array, initial_spot_price(),
                               volatility(), sqrt_dt(), risk_free_rate(), contract_strike(), dt(), payoffs.data());
        return payoffs;
    }

    // GPU twin of simulate_trajectory_cpu on the SAME normals (the reference has the kernels,
    // inc/trajectories.cuh:14-52, but never calls them): per-path call payoffs.
    std::vector<float> simulate_trajectory_gpu()
    {
        if (!d_random_array) initialize_random_array();
        std::vector<float> payoffs(n_trajectories, 0.0f);
        OptionData od = option_data();
        mcb_engine *e = mcb_compat_engine();
        std::vector<float> dev_out(0);
        float *d_pay = nullptr;
        testCUDA(cudaMalloc(&d_pay, n_trajectories * sizeof(float)));
        if (!e || mcb_price_from_normals(e, mcb_compat_cast(od), d_random_array, n_trajectories, (int)n_steps, d_pay,
                                         MCB_DEVICE) != MCB_OK) {
            fprintf(stderr, "mcb200: %s\n", mcb_last_error());
            exit(EXIT_FAILURE);
        }
        testCUDA(cudaMemcpy(payoffs.data(), d_pay, n_trajectories * sizeof(float), cudaMemcpyDeviceToHost));
        cudaFree(d_pay);
        return payoffs;
    }

    // Full trajectories, path-major out[p*n_steps + i] = S(t_{i+1}) (inc/testing.cuh:281-326).
    std::vector<float> simulate_outer_trajectories(size_t n_threads_per_block, uint64_t seed)
    {
        const size_t blocks = (n_trajectories + n_threads_per_block - 1) / n_threads_per_block;
        std::cout << "====================================================================\n";
        std::cout << "Going to compute outer trajectories...\n";
        std::cout << "Number of threads per block: " << n_threads_per_block << "\n";
        std::cout << "Number of trajectories: " << n_trajectories << "\n";
        std::cout << "Number of blocks: " << blocks << "\n";
        std::cout << "Number of steps: " << n_steps << "\n";
        std::cout << "====================================================================\n";
        std::vector<float> out(length());
        OptionData od = option_data();
        mcb_engine *e = mcb_compat_engine();
        if (!e || mcb_simulate_trajectories(e, mcb_compat_cast(od), 0, n_trajectories, seed, out.data(), nullptr,
                                            MCB_HOST) != MCB_OK) {
            fprintf(stderr, "mcb200: %s\n", mcb_last_error());
            exit(EXIT_FAILURE);
        }
        return out;
    }

    size_t length() { return n_steps * n_trajectories; }

    void initialize_random_array(size_t seed = 1234ULL)
    {
        if (d_random_array) testCUDA(cudaFree(d_random_array));
        if (h_random_array) free(h_random_array);
        init_random_array(&d_random_array, &h_random_array, length(), (long)seed);
    }

    float &volatility() { return sigma; }
    float &risk_free_rate() { return r; }
    float &initial_spot_price() { return x_0; }
    float &contract_strike() { return K; }
    float &contract_maturity() { return T; }
    float &barrier() { return B; }
    float dt() { return contract_maturity() / n_steps; }
    float sqrt_dt() { return sqrt(dt()); }

    OptionData option_data()
    {
        OptionData od{};
        od.S0 = x_0; od.T = T; od.K = K; od.r = r; od.v = sigma; od.B = B;
        od.P1 = (int)P1; od.P2 = (int)P2;
        od.N_PATHS = (int)n_trajectories; od.N_PATHS_INNER = 0; od.N_STEPS = (int)n_steps;
        od.step = dt();
        return od;
    }

    float sigma;  // volatility
    float r;      // risk-free rate
    float x_0;    // initial spot price
    float K;      // contract strike
    float T;      // contract maturity
    float B;      // barrier
    float P1;
    float P2;
};
