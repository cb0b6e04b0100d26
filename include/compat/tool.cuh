// compat/tool.cuh -- drop-in for the reference's inc/tool.cuh over libmcb200.so.
//
// Same names a caller of the reference uses (inc/tool.cuh): struct OptionData (:13-26),
// printOptionData (:29-44), getDeviceProperty (:56-88), testCUDA (:92-100),
// simulateOptionPriceCPU / simulateBulletOptionPriceCPU (:104-173, the reference's explicitly-CPU
// twins behind wrapper_cpu_*), get_max_blocks (:176-188), isPow2 / nextPow2 (:200-210).
// setup_kernel (:192-195) has no counterpart: the engine's Philox stream is stateless.
// Written from scratch; nothing here is on the GPU hot path.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <string>

#include <cuda_runtime.h>

#include "../mcb200.h"

using namespace std;  // the reference's headers export this, and hello.cu relies on it (cout, endl)

// Byte-identical to mcb_option_data (static_assert below) and to the reference's struct.
struct OptionData {
    float S0, T, K, r, v, B;
    int P1, P2, N_PATHS, N_PATHS_INNER, N_STEPS;
    float step;
};
static_assert(sizeof(OptionData) == sizeof(mcb_option_data), "OptionData must stay 48 bytes, 12 fields");

inline const mcb_option_data *mcb_compat_cast(const OptionData &o)
{
    return reinterpret_cast<const mcb_option_data *>(&o);
}

inline void printOptionData(OptionData od)
{
    const char *names[] = {"S0", "T", "K", "r", "v", "B"};
    const float reals[] = {od.S0, od.T, od.K, od.r, od.v, od.B};
    cout << endl;
    for (int i = 0; i < 6; ++i) cout << names[i] << " : " << reals[i] << endl;
    cout << "P1 : " << od.P1 << endl << "P2 : " << od.P2 << endl;
    cout << "N_PATHS : " << od.N_PATHS << endl << "N_PATHS_INNER : " << od.N_PATHS_INNER << endl;
    cout << "N_STEPS : " << od.N_STEPS << endl << "step : " << od.step << endl << endl;
}

#define CHECK_MALLOC(ptr)                                                                            \
    do {                                                                                             \
        if ((ptr) == NULL) {                                                                         \
            fprintf(stderr, "Memory allocation failed for %s at %s:%d\n", #ptr, __FILE__, __LINE__); \
            exit(EXIT_FAILURE);                                                                      \
        }                                                                                            \
    } while (0)

// Same contract as the reference's macro: print and exit on a CUDA error (callers' own code only;
// the library itself never exits).
inline void mcb_compat_test_cuda(cudaError_t error, const char *file, int line)
{
    if (error == cudaSuccess) return;
    fprintf(stderr, "There is an error in file %s at line %d: %s\n", file, line, cudaGetErrorString(error));
    exit(EXIT_FAILURE);
}
#define testCUDA(error) (mcb_compat_test_cuda(error, __FILE__, __LINE__))

// One process-wide engine.  The reference always runs on the implicit device 0 (inc/wrappers.cuh:33-57);
// so does this shim by default.  MCB200_DEVICES opts the SAME unmodified caller (hello.cu, testing.cu) into
// every GPU of the box -- "all", a count ("8") or a device list ("0,1,2,3") -- through
// mcb_engine_create_multi: the wrappers then shard internally and return the same bits.
inline mcb_engine *mcb_compat_engine()
{
    static mcb_engine *engine = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        int devices[MCB_MAX_PEERS] = {0};
        int n = 1;
        if (const char *env = getenv("MCB200_DEVICES")) {
            int count = 0;
            cudaGetDeviceCount(&count);
            const std::string spec(env);
            if (spec == "all") {
                n = count;
            } else if (spec.find(',') == std::string::npos) {
                n = atoi(env);
            } else {
                n = 0;
                size_t pos = 0;
                while (pos <= spec.size() && n < MCB_MAX_PEERS) {
                    const size_t comma = spec.find(',', pos);
                    devices[n++] = atoi(spec.substr(pos, comma == std::string::npos ? comma : comma - pos).c_str());
                    if (comma == std::string::npos) break;
                    pos = comma + 1;
                }
            }
            if (n < 1) n = 1;
            if (n > MCB_MAX_PEERS) n = MCB_MAX_PEERS;
            if (spec.find(',') == std::string::npos)
                for (int i = 0; i < n; ++i) devices[i] = i;
        }
        if (mcb_engine_create_multi(devices, n, &engine) != MCB_OK) engine = nullptr;  // message stays in mcb_last_error()
    }
    return engine;
}

inline void getDeviceProperty()
{
    mcb_device_info info;
    mcb_engine *e = mcb_compat_engine();
    if (!e || mcb_get_device_info(e, &info) != MCB_OK) {
        cout << "No usable CUDA device: " << mcb_last_error() << endl;
        return;
    }
    cout << "Device : " << info.name << " (sm_" << info.cc_major << info.cc_minor << ")" << endl;
    cout << "Multiprocessors : " << info.sm_count << endl;
    cout << "Clock rate : " << info.clock_khz / 1000 << " MHz" << endl;
    cout << "Global memory : " << (double)info.total_mem / (1024.0 * 1024.0 * 1024.0) << " GiB" << endl << endl;
}

inline int get_max_blocks(int threadsPerBlock)
{
    mcb_device_info info;
    mcb_engine *e = mcb_compat_engine();
    if (!e || mcb_get_device_info(e, &info) != MCB_OK || threadsPerBlock <= 0) return 0;
    return info.sm_count * (2048 / threadsPerBlock > 0 ? 2048 / threadsPerBlock : 1);
}

inline bool isPow2(unsigned int x) { return x != 0 && (x & (x - 1)) == 0; }

inline unsigned int nextPow2(unsigned int x)
{
    unsigned int p = 1;
    while (p < x && p != 0) p <<= 1;
    return p;
}

// ---- the reference's CPU twins (explicitly-CPU API, NOT a fallback for the GPU entry points) ----
// Same estimator, same unseeded std::mt19937 + normal_distribution<float> draw as inc/tool.cuh:116-117,
// but the payoff sum is kept in double so the price stays valid past ~1e7 paths.
inline void simulateOptionPriceCPU(float *optionPriceCPU, OptionData o)
{
    mt19937 gen(random_device{}());
    normal_distribution<float> gauss(0.0f, 1.0f);
    const float drift = (o.r - 0.5f * o.v * o.v) * o.T, vol = o.v * sqrtf(o.T);
    double acc = 0.0;
    for (int p = 0; p < o.N_PATHS; ++p) {
        const float terminal = o.S0 * expf(drift + vol * gauss(gen));
        if (terminal > o.K) acc += (double)(terminal - o.K);
    }
    *optionPriceCPU = (float)(exp(-(double)o.r * o.T) * acc / (double)o.N_PATHS);
}

inline void simulateBulletOptionPriceCPU(float *optionPriceCPU, OptionData o)
{
    mt19937 gen(random_device{}());
    normal_distribution<float> gauss(0.0f, 1.0f);
    const float drift = (o.r - 0.5f * o.v * o.v) * o.step, vol = o.v * sqrtf(o.step);
    const float log_barrier = o.B > 0.0f ? logf(o.B) : -INFINITY;
    double acc = 0.0;
    for (int p = 0; p < o.N_PATHS; ++p) {
        float log_s = logf(o.S0);
        int below = 0;
        for (int k = 0; k < o.N_STEPS; ++k) {
            log_s += drift + vol * gauss(gen);
            below += log_s < log_barrier;
        }
        const float terminal = expf(log_s);
        if (below >= o.P1 && below <= o.P2 && terminal > o.K) acc += (double)(terminal - o.K);
    }
    *optionPriceCPU = (float)(exp(-(double)o.r * o.T) * acc / (double)o.N_PATHS);
}
