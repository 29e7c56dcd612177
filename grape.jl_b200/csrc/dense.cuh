// dense.cuh -- placeholder (filled in below in this round)
#pragma once
#include "common.cuh"
#include "../../include/grape_b200.h"
#include <string>
#include <vector>
struct DensePlan { int dummy; };
inline int dense_setup(DensePlan&, DevP&, const grape_b200_problem*, std::vector<void*>&, std::string& e) { e = "dense path not built"; return GRAPE_B200_EINVAL; }
inline void dense_destroy(DensePlan&) {}
inline void dense_run_forward(DensePlan&, const DevP&, cudaStream_t, int64_t&) {}
inline void dense_run_backward(DensePlan&, const DevP&, const cplx*, cudaStream_t, int64_t&) {}
inline void dense_gather_final(DensePlan&, const DevP&, cplx*, cudaStream_t, int64_t&) {}
inline void dense_gather_states(DensePlan&, const DevP&, int, cplx*, cudaStream_t, int64_t&) {}
