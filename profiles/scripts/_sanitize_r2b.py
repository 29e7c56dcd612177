"""GPU session helper (not a test): compute-sanitizer (memcheck) driver for the kernels added late in round 2 -- the
economised-polynomial tables (dense chains strip / tiled / concurrent, real-symmetric small path with the 128-register
gradient kernel, out-of-line orders 7 / 8 and sub-steps), and the scan schedule of the sub-warp path (radix 2 and 4)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grape.jl_b200 as gb  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402
from grape.jl_b200.engine import GrapeEngine  # noqa: E402


def run(p, eps, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        e = GrapeEngine(p)
    finally:
        for k in env:
            os.environ.pop(k, None)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    e.evaluate_functional(eps)
    if p.N <= 32:
        e.stored_states(0)
    J2 = e.evaluate_gradient(G, eps)
    assert J == J2
    e.close()
    return J


if __name__ == "__main__":
    p, eps = configs.c3_ensemble(n_delta=5, n_amp=8, NT=300)
    for scale in (1.0, 4.0, 12.0, 60.0):     # orders 5..6 inline, 7..8 out of line, sub-stepped
        for env in (dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_SEG_SCAN=1), dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_SEG_SCAN=0),
                    dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_ECON=0)):
            print("c3", scale, env, run(p, eps * scale, **env))
    for d2 in (0, 1):
        pd, ed = configs.c5_dense1024(N=64, K=16, NT=5)
        print("dense econ", d2, run(pd, ed, GRAPE_B200_DENSE2=d2))
    pd, ed = configs.c4_dense450(N=64, K=16, NT=5)
    print("dense econ concurrent", run(pd, ed, GRAPE_B200_DENSE2=0))
    for N, NT, radix in ((6, 70, 4), (6, 70, 2), (12, 40, 4), (20, 40, 2), (32, 33, 2)):
        pw, ew = configs.random_problem(K=3, N=N, L=2, NT=NT, seed=5, hermitian=True, functional=gb.SM, G=2)
        pw.tlist[:] = pw.tlist * (0.6 / np.sqrt(N))
        print("warp scan", N, NT, radix, run(pw, ew, GRAPE_B200_WSEG_RADIX=radix), run(pw, ew, GRAPE_B200_WSEG_RADIX=radix, GRAPE_B200_SEG_S=2))
    print("SANITIZE_R2B_DONE")
