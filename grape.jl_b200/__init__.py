"""grape.jl_b200 -- B200-native GRAPE gradient engine behind the reference's
`GRAPE.optimize(trajectories, tlist; prop_method, J_T, chi, J_a, ...)` API.

Host-side mirror of the reference interface for the hot path only
(reference src/optimize.jl:63-144, 696-768, 824-1014; src/workspace.jl:78-362).
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI in
include/grape_b200.h; there is no CPU fallback: using the engine without the
built shared library raises."""
from .problem import (GrapeProblem, SM, RE, SS, HOST, GRADGEN, TAYLOR, JA_NONE, JA_FLUENCE,
                      GB_NONE, GB_QUADFORM, PATH_AUTO, PATH_SMALL, PATH_WARP, PATH_DENSE, PATH_SMALL_CHAIN, PATH_WARP_CHAIN)
from . import configs

__all__ = ["GrapeProblem", "configs", "SM", "RE", "SS", "HOST", "GRADGEN", "TAYLOR",
           "JA_NONE", "JA_FLUENCE", "GB_NONE", "GB_QUADFORM"]
