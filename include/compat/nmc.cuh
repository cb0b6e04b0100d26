// compat/nmc.cuh -- name kept for callers that include inc/nmc.cuh directly.  The three nested-MC
// kernels (inc/nmc.cuh:12-386) are replaced by nested_kernel inside libmcb200.so, reached through
// the wrapper_gpu_bullet_option_nmc_* shims (compat/wrappers.cuh) or mcb_nested_monte_carlo.
#pragma once
#include "tool.cuh"
