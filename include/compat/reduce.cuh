// compat/reduce.cuh -- name kept for callers that include inc/reduce.cuh directly.  The reference's
// reduce3..6 kernels (inc/reduce.cuh:9-227) are replaced by the engine's deterministic
// warp-shuffle + shared-memory tree inside libmcb200.so; their per-block index ranges are
// available through mcb_reduce_blocks (see compat/testing.cuh, Simulation::test_reduction).
#pragma once
#include "tool.cuh"
