"""ctypes binding of the C-ABI shared library (include/grape_b200.h).

The library is built in-tree by `build.py` (nvcc, sm_100a) as
`grape.jl_b200/libgrape_b200.so`.  There is no fallback: if the library is
missing or cannot be loaded, `load()` raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrape_b200.so")

ABI_VERSION = 1

ERROR_NAMES = {0: "OK", 1: "EINVAL", 2: "ECUDA", 3: "ECHINORM", 4: "ETAYLOR", 5: "ENOCONTROLS",
               6: "ESTATE", 7: "ENCCL"}


class ProblemDesc(C.Structure):
    """Mirror of `grape_b200_problem`."""
    _fields_ = [
        ("abi_version", C.c_int32), ("K", C.c_int32), ("N", C.c_int32), ("L", C.c_int32),
        ("NT", C.c_int32), ("G", C.c_int32), ("K_global", C.c_int32), ("device", C.c_int32),
        ("tlist", C.POINTER(C.c_double)), ("gen_of_traj", C.POINTER(C.c_int32)),
        ("H0", C.POINTER(C.c_double)), ("Hc", C.POINTER(C.c_double)),
        ("shape", C.POINTER(C.c_double)), ("psi0", C.POINTER(C.c_double)),
        ("tgt", C.POINTER(C.c_double)), ("weights", C.POINTER(C.c_double)),
        ("functional", C.c_int32), ("gradient_method", C.c_int32), ("ja_kind", C.c_int32),
        ("gb_kind", C.c_int32), ("lambda_a", C.c_double), ("lambda_b", C.c_double),
        ("gb_D", C.POINTER(C.c_double)), ("gb_nD", C.c_int32), ("taylor_max_order", C.c_int32),
        ("taylor_tolerance", C.c_double), ("taylor_check_convergence", C.c_int32),
        ("path", C.c_int32), ("chi_min_norm", C.c_double),
    ]


# every symbol include/grape_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
SYMBOLS = {
    "grape_b200_abi_version": (C.c_int, []),
    "grape_b200_create": (C.c_int, [C.POINTER(ProblemDesc), C.POINTER(_P)]),
    "grape_b200_destroy": (None, [_P]),
    "grape_b200_last_error": (C.c_char_p, [_P]),
    "grape_b200_eval_f": (C.c_int, [_P, _D, _D, _D]),
    "grape_b200_eval_fg": (C.c_int, [_P, _D, _D, _D, _D, _D, _D]),
    "grape_b200_eval_f_amplitudes": (C.c_int, [_P, _D, _D, _D]),
    "grape_b200_eval_fg_amplitudes": (C.c_int, [_P, _D, _D, _D, _D, _D]),
    "grape_b200_forward": (C.c_int, [_P, _D, _D, _D]),
    "grape_b200_backward": (C.c_int, [_P, _D, _D, _D, _D]),
    "grape_b200_backward_chi": (C.c_int, [_P, _D, _D, _D, _D]),
    "grape_b200_get_final_states": (C.c_int, [_P, _D]),
    "grape_b200_get_stored_states": (C.c_int, [_P, C.c_int32, _D]),
    "grape_b200_get_chi_states": (C.c_int, [_P, _D, _D]),
    "grape_b200_get_tau_grads": (C.c_int, [_P, C.c_int32, _D]),
    "grape_b200_get_timings": (C.c_int, [_P, _D]),
    "grape_b200_set_profiling": (C.c_int, [_P, C.c_int32]),
    "grape_b200_eval_fg_device": (C.c_int, [_P, _P, _P, _P]),
    "grape_b200_enqueue_forward": (C.c_int, [_P, _P]),
    "grape_b200_enqueue_backward": (C.c_int, [_P]),
    "grape_b200_enqueue_combine": (C.c_int, [_P]),
    "grape_b200_finish": (C.c_int, [_P]),
    "grape_b200_device_ptr": (_P, [_P, C.c_int32]),
    "grape_b200_stream": (_P, [_P]),
    "grape_b200_launch_count": (C.c_int64, [_P]),
    "grape_b200_gradient_form": (C.c_int, [_P]),
    "grape_b200_small_schedule": (C.c_int, [_P]),
    "grape_b200_dense_concurrent": (C.c_int, [_P]),
    "grape_b200_dense_orders": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "grape_b200_econ_table": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "grape_b200_xchg_init": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "grape_b200_xchg_attach": (C.c_int, [_P, _P]),
    "grape_b200_xchg_detach": (C.c_int, [_P]),
    "grape_b200_multi_create": (C.c_int, [C.POINTER(ProblemDesc), C.POINTER(C.c_int32), C.c_int32, C.POINTER(_P)]),
    "grape_b200_multi_destroy": (None, [_P]),
    "grape_b200_multi_last_error": (C.c_char_p, [_P]),
    "grape_b200_multi_eval_f": (C.c_int, [_P, _D, _D, _D]),
    "grape_b200_multi_eval_fg": (C.c_int, [_P, _D, _D, _D, _D, _D, _D]),
    "grape_b200_multi_get_final_states": (C.c_int, [_P, _D]),
    "grape_b200_multi_size": (C.c_int32, [_P]),
    "grape_b200_multi_shard": (_P, [_P, C.c_int32, C.POINTER(C.c_int32)]),
}
IPC_HANDLE_BYTES = 64

_lib = None


def load():
    """Load the shared library and bind every declared symbol. Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). grape.jl_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.grape_b200_abi_version() != ABI_VERSION:
        raise RuntimeError("libgrape_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib
