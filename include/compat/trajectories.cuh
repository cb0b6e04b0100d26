// compat/trajectories.cuh -- keeps the one symbol of inc/trajectories.cuh a caller touches:
// `__constant__ OptionData d_OptionData` (inc/trajectories.cuh:12), which hello.cu:22 uploads with
// cudaMemcpyToSymbol before calling the wrappers.  The new engine passes parameters as kernel
// arguments, so the symbol is accepted and ignored; the kernels themselves live in libmcb200.so.
#pragma once
#include "BlackandScholes.hpp"
#include "tool.cuh"

#ifdef __CUDACC__
__constant__ OptionData d_OptionData;
#endif
