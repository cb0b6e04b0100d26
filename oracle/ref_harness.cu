// ref_harness.cu -- ORACLE / TEST INFRASTRUCTURE ONLY.
// Thin extern "C" doors onto the UNMODIFIED reference sources, which are compiled where
// they lie (-I/root/reference/inc, see oracle/Makefile); nothing is copied into this repo.
// Output goes to oracle/_ref/libref_cpu.so (git-ignored, travels to the GPU box).
//   simulateOptionPriceCPU        inc/tool.cuh:104-130
//   simulateBulletOptionPriceCPU  inc/tool.cuh:133-173
//   black_scholes_CPU, CND        inc/BlackandScholes.hpp:8-43
#include <cstdint>
#include "tool.cuh"
#include "BlackandScholes.hpp"

extern "C" {

int ref_sizeof_option_data() { return (int)sizeof(OptionData); }

float ref_vanilla_cpu(const OptionData *o)
{
    float price = 0.0f;
    simulateOptionPriceCPU(&price, *o);
    return price;
}

float ref_bullet_cpu(const OptionData *o)
{
    float price = 0.0f;
    simulateBulletOptionPriceCPU(&price, *o);
    return price;
}

// The reference's single float accumulator is only accurate to ~1e6-1e7 paths per call
// (SURVEY.md row a5), so long runs are cut into calls of `chunk` paths; the mean of the
// chunk prices is taken in double.  Returns the mean price; *paths_done = paths simulated.
double ref_vanilla_cpu_chunked(const OptionData *o, uint64_t total_paths, int chunk, uint64_t *paths_done)
{
    OptionData c = *o;
    double acc = 0.0;
    uint64_t done = 0, calls = 0;
    while (done < total_paths) {
        uint64_t n = total_paths - done < (uint64_t)chunk ? total_paths - done : (uint64_t)chunk;
        c.N_PATHS = (int)n;
        float price = 0.0f;
        simulateOptionPriceCPU(&price, c);
        acc += (double)price * (double)n;
        done += n;
        ++calls;
    }
    if (paths_done) *paths_done = done;
    return done ? acc / (double)done : 0.0;
}

double ref_bullet_cpu_chunked(const OptionData *o, uint64_t total_paths, int chunk, uint64_t *paths_done)
{
    OptionData c = *o;
    double acc = 0.0;
    uint64_t done = 0;
    while (done < total_paths) {
        uint64_t n = total_paths - done < (uint64_t)chunk ? total_paths - done : (uint64_t)chunk;
        c.N_PATHS = (int)n;
        float price = 0.0f;
        simulateBulletOptionPriceCPU(&price, c);
        acc += (double)price * (double)n;
        done += n;
    }
    if (paths_done) *paths_done = done;
    return done ? acc / (double)done : 0.0;
}

float ref_black_scholes(float S0, float K, float T, float r, float v)
{
    float call = 0.0f;
    black_scholes_CPU(call, S0, K, T, r, v);
    return call;
}

float ref_cnd(float x) { return CND(x); }

}  // extern "C"
