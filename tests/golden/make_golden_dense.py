"""Generates the FULL-SIZE golden vectors of BASELINE configs[3] / [4]:

    python tests/golden/make_golden_dense.py [c4] [c5]

    c4_dense450_full.npz    K=16, N=450,  L=2, NT=5000, J_T_sm
    c5_dense1024_full.npz   K=64, N=1024, L=2, NT=1000, J_T_sm + J_a fluence + quadratic g_b

PROVENANCE: produced offline (about 10 / 25 CPU-minutes on 8 cores) by the spectral CPU oracle
oracle/dense_oracle.py -- per step one Hermitian eigendecomposition, exact propagator and exact
Frechet derivative in the eigenbasis -- which restates reference src/optimize.jl:696-768, 824-911,
574-584, 1002-1011 and is pinned to the literal restatement (oracle/grape_oracle.py) at 1e-12 on
reduced sizes (tests/test_dense_oracle.py).  The algorithm shares nothing with the CUDA path's
Taylor / Krylov series, so a chained error over 5000 steps cannot cancel.  The operators are NOT
stored (3 x N^2 complex: 10 / 50 MB); they are regenerated from the seeded generators of
grape.jl_b200/configs.py and a fingerprint is stored instead."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from grape.jl_b200 import configs  # noqa: E402
from oracle import dense_oracle as do  # noqa: E402


def dump(name, p, eps):
    t0 = time.time()
    r = do.evaluate_gradient(p, eps, keep_eig=True, progress=250)
    fp = np.array([np.sum(p.H0).real, np.sum(np.abs(p.Hc)), np.sum(p.tgt).imag])
    np.savez_compressed(
        os.path.join(HERE, "dense_full", name + ".npz"), pulsevals=eps, J=r["J"], J_parts=r["J_parts"], tau=r["tau"], G=r["G"],
        grad_J_Tb=r["grad_J_Tb"], grad_J_a=r["grad_J_a"], chi_norms=r["chi_norms"],
        final_state_norms=np.linalg.norm(r["final_states"], axis=1),
        final_states_head=r["final_states"][:, :8].copy(), fingerprint=fp,
        dims=np.array([p.K, p.N, p.L, p.NT]))
    print(f"{name}: J={r['J']:.15f} |G|={np.linalg.norm(r['G']):.15e}  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c5"]
    if "c4" in which:
        dump("c4_dense450_full", *configs.c4_dense450())
    if "c5" in which:
        dump("c5_dense1024_full", *configs.c5_dense1024())
