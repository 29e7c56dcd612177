// warp_n.cuh -- placeholder (filled in below in this round)
#pragma once
#include "common.cuh"
#include "../../include/grape_b200.h"
#include <string>
#include <vector>
constexpr int WARP_MAX_N = 64;
struct WarpPlan { int W; };
inline int warp_setup(WarpPlan&, DevP&, const grape_b200_problem*, std::vector<void*>&, std::string& e) { e = "warp path not built"; return GRAPE_B200_EINVAL; }
inline void warp_run_formU(WarpPlan&, const DevP&, cudaStream_t, int64_t&) {}
inline void warp_run_forward(WarpPlan&, const DevP&, cudaStream_t, int64_t&) {}
inline void warp_run_backward(WarpPlan&, const DevP&, const cplx*, cudaStream_t, int64_t&) {}
inline void warp_run_gradient(WarpPlan&, const DevP&, cudaStream_t, int64_t&) {}
inline void warp_gather_final(WarpPlan&, const DevP&, cplx*, cudaStream_t, int64_t&) {}
inline void warp_gather_states(WarpPlan&, const DevP&, int, cplx*, cudaStream_t, int64_t&) {}
