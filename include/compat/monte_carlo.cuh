// compat/monte_carlo.cuh -- the reference's include aggregator (inc/monte_carlo.cuh:3-8): a caller
// that does `#include "monte_carlo.cuh"` (hello.cu:1) gets the same symbols, backed by libmcb200.so.
#pragma once
#include "nmc.cuh"
#include "reduce.cuh"
#include "testing.cuh"
#include "tool.cuh"
#include "trajectories.cuh"
#include "wrappers.cuh"
#include "option_price.hpp"
