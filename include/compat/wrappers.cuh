// compat/wrappers.cuh -- the reference's call surface (inc/wrappers.cuh:10-340) as inline shims
// over the C-ABI of libmcb200.so.  Same names, same argument meaning, same float return, same
// stdout side effect; `threadsPerBlock` / `number_of_blocks` are accepted and ignored (results do
// not depend on launch geometry).  A failing engine call maps to the reference's `return -1`
// (inc/wrappers.cuh:74-78).  Seeds are the reference's hard-coded 1234 / 1235
// (inc/wrappers.cuh:41,151,163).
#pragma once
#include "tool.cuh"
#include "trajectories.cuh"

inline float mcb_compat_report(const char *label, int status, double value)
{
    if (status != MCB_OK) {
        fprintf(stderr, "mcb200: %s\n", mcb_last_error());
        return -1.0f;
    }
    cout << label << (float)value << endl << endl;
    return (float)value;
}

inline float wrapper_cpu_option_vanilla(OptionData option_data, int /*threadsPerBlock*/)
{
    float price = 0.0f;
    simulateOptionPriceCPU(&price, option_data);
    cout << endl << "Average CPU Vanilla Option: " << price << endl << endl;
    return price;
}

inline float wrapper_cpu_bullet_option(OptionData option_data, int /*threadsPerBlock*/)
{
    float price = 0.0f;
    simulateBulletOptionPriceCPU(&price, option_data);
    cout << endl << "Monte Carlo CPU Bullet Option Price : " << price << endl << endl;
    return price;
}

inline float wrapper_gpu_option_vanilla(OptionData option_data, int /*threadsPerBlock*/)
{
    mcb_result res{};
    mcb_engine *e = mcb_compat_engine();
    const int st = e ? mcb_price_european(e, mcb_compat_cast(option_data), 0, 1234, MCB_CALL, &res) : MCB_ERR_NO_DEVICE;
    return mcb_compat_report("Average GPU : ", st, res.price);
}

inline float wrapper_gpu_bullet_option(OptionData option_data, int /*threadsPerBlock*/)
{
    mcb_result res{};
    mcb_engine *e = mcb_compat_engine();
    const int st = e ? mcb_price_bullet(e, mcb_compat_cast(option_data), 0, 1234, 0, 0.0f, 0, &res) : MCB_ERR_NO_DEVICE;
    return mcb_compat_report("Average GPU bullet option : ", st, res.price);
}

// The engine has no atomics: same estimator and same bits as wrapper_gpu_bullet_option.
inline float wrapper_gpu_bullet_option_atomic(OptionData option_data, int /*threadsPerBlock*/)
{
    mcb_result res{};
    mcb_engine *e = mcb_compat_engine();
    const int st = e ? mcb_price_bullet(e, mcb_compat_cast(option_data), 0, 1234, 0, 0.0f, 0, &res) : MCB_ERR_NO_DEVICE;
    return mcb_compat_report("Average GPU bullet option atomic : ", st, res.price);
}

// The three nested-MC wrappers return the reference's diagnostic scalar: the average of F over
// N_PATHS*N_STEPS + 1 slots (inc/wrappers.cuh:134,185-189).  F itself is available through
// mcb_nested_monte_carlo.
inline float mcb_compat_nested(OptionData option_data, const char *label)
{
    mcb_engine *e = mcb_compat_engine();
    double mean = 0.0;
    int st = MCB_ERR_NO_DEVICE;
    if (e && option_data.N_PATHS > 0 && option_data.N_STEPS > 0) {
        const size_t n = (size_t)option_data.N_PATHS * (size_t)option_data.N_STEPS;
        float *F = (float *)malloc(n * sizeof(float));
        CHECK_MALLOC(F);
        st = mcb_nested_monte_carlo(e, mcb_compat_cast(option_data), 0, (uint64_t)option_data.N_PATHS, 1234, 1235,
                                    MCB_DISCOUNT_COMPAT, F, nullptr, nullptr, MCB_HOST, &mean);
        free(F);
    }
    return mcb_compat_report(label, st, mean);
}

inline float wrapper_gpu_bullet_option_nmc_one_point_one_block(OptionData option_data, int /*threadsPerBlock*/,
                                                               int /*number_of_blocks*/)
{
    return mcb_compat_nested(option_data, "Average GPU bullet option nmc one point per block : ");
}

inline float wrapper_gpu_bullet_option_nmc_one_kernel(OptionData option_data, int /*threadsPerBlock*/,
                                                      int /*number_of_blocks*/)
{
    return mcb_compat_nested(option_data, "Average GPU bullet option nmc one kernel : ");
}

inline float wrapper_gpu_bullet_option_nmc_optimal(OptionData option_data, int /*threadsPerBlock*/,
                                                   int /*number_of_blocks*/)
{
    return mcb_compat_nested(option_data, "Average GPU bullet option nmc optimal : ");
}
