"""GPU session helper (not a test): small driver for compute-sanitizer (memcheck / racecheck) over the kernels added in
round 2 -- staged real-symmetric kernels with in-kernel sub-stepping, scan schedule (formscan / tau / bounds and the
Q-derived prologue), one-pass ring chain, fused finalize, exchange kernels (three shards on one device), amplitude mode,
concurrent dense chains with TMA strip prefetch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grape.jl_b200 as gb  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402
from grape.jl_b200.engine import GrapeEngine, MultiGrapeEngine  # noqa: E402


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
    if p.N <= 32:
        e.tau_grads(0)
    e.evaluate_functional(eps)
    e.stored_states(0)
    J2 = e.evaluate_gradient(G, eps)
    assert J == J2
    e.close()
    return J


if __name__ == "__main__":
    p, eps = configs.c3_ensemble(n_delta=5, n_amp=8, NT=300)
    for env in (dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_SEG_SCAN=1), dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_SEG_SCAN=0),
                dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_SEG_SCAN=0, GRAPE_B200_CHAIN_DUAL=0),
                dict(GRAPE_B200_FORCE_FORMSEG=1, GRAPE_B200_SEG_S=3)):
        print("c3", env, run(p, eps, **env))
        print("c3 x30 (sub-steps)", run(p, eps * 30.0, **env))
    pm, epsm = configs.c3_ensemble(n_delta=5, n_amp=7, NT=90, functional=gb.SM, ja_kind=1, lambda_a=0.05)
    m = MultiGrapeEngine(pm, [0, 0, 0])
    G = np.zeros_like(epsm)
    print("multi", m.evaluate_gradient(G, epsm), m.evaluate_functional(epsm), m.evaluate_gradient(G, epsm))
    m.close()
    p1, e1 = configs.c3_ensemble(n_delta=4, n_amp=4, NT=60)
    e = GrapeEngine(p1)
    a = e1.copy()
    Gs = np.zeros_like(a)
    print("amplitudes", e.evaluate_gradient_amplitudes(Gs, a, np.ones_like(a)), e.evaluate_functional_amplitudes(a))
    e.close()
    for conc, tma in ((1, 1), (0, 0)):
        pd, ed = configs.c4_dense450(N=64, K=16, NT=5)
        print("dense", conc, tma, run(pd, ed, GRAPE_B200_DENSE2=0, GRAPE_B200_DENSE_CONCURRENT=conc, GRAPE_B200_DENSE_TMA=tma))
    print("SANITIZE_R2_DONE")
