"""Generates tests/golden/*.npz -- golden input/output vectors of the hot path.

    python tests/golden/make_golden.py

PROVENANCE (read this): the reference (GRAPE.jl) is Julia and cannot run in
this image, and its own tests store no J_T / gradient / pulse vectors
(SURVEY.md 8c).  These vectors are therefore produced by the CPU oracle
(oracle/grape_oracle.py: NumPy + SciPy `expm`, the dense per-trajectory
ExpProp + GradGenerator algorithm of reference src/optimize.jl:696-768,
824-1014), *after* that oracle has been pinned by the reference's RNG-free
known answers (tests/test_oracle_known_answers.py).  They freeze the oracle's
outputs so that (a) the C restatement, (b) the CUDA path on the GPU box and (c)
any later change of the oracle itself are compared with committed numbers, and
so that anybody with Julia can diff real GRAPE.jl output against the same
inputs (every case stores the complete problem definition).

Each file: problem arrays (tlist, H0, Hc, psi0, tgt, gen_of_traj, shape,
weights, gb_D, scalars) + pulsevals + J, J_parts, tau, G, grad_J_Tb, grad_J_a,
final_states."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from grape.jl_b200 import configs  # noqa: E402
from oracle import grape_oracle as go  # noqa: E402

D3 = np.diag([0.0, 1.0, 0.0]).astype(complex)


def cases():
    yield "c1_readme_full", configs.c1_readme()
    yield "tls_fixture_full", configs.tls_fixture()
    yield "c2_transmon_nt200", configs.c2_transmon(NT=200)
    yield "c3_ensemble_4x4_nt100_ss", configs.c3_ensemble(n_delta=4, n_amp=4, NT=100)
    yield "c3_ensemble_3x5_nt60_sm_gb", configs.c3_ensemble(
        n_delta=3, n_amp=5, NT=60, functional=0, gb_kind=1, gb_D=D3, lambda_b=0.4, ja_kind=1, lambda_a=0.05)
    yield "c4_dense_n48_k5_nt12", configs.c4_dense450(N=48, K=5, NT=12)
    yield "c5_dense_n40_k8_nt10_costs", configs.c5_dense1024(N=40, K=8, NT=10)
    yield "random_nonherm_shaped_weighted", configs.random_problem(
        K=6, N=4, L=3, NT=23, G=2, seed=77, hermitian=False, shaped=True, functional=1,
        weights=np.linspace(0.5, 1.5, 6))
    yield "random_n7_taylor", configs.random_problem(K=3, N=7, L=2, NT=15, seed=78, gradient_method=1)


def dump(name, p, eps):
    r = go.evaluate_gradient(go.from_problem(p), eps)
    out = dict(
        tlist=p.tlist, H0=p.H0, Hc=p.Hc, psi0=p.psi0, tgt=p.tgt, gen_of_traj=p.gen_of_traj,
        shape=np.zeros(0) if p.shape is None else p.shape,
        weights=np.zeros(0) if p.weights is None else p.weights,
        gb_D=np.zeros(0) if p.gb_D is None else p.gb_D,
        scalars=np.array([p.functional, p.gradient_method, p.ja_kind, p.gb_kind], dtype=np.int64),
        lambdas=np.array([p.lambda_a, p.lambda_b]),
        pulsevals=eps, J=r["J"], J_parts=r["J_parts"], tau=r["tau"], G=r["G"],
        grad_J_Tb=r["grad_J_Tb"], grad_J_a=r["grad_J_a"], final_states=r["final_states"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: J={r['J']:.15f} |G|={np.linalg.norm(r['G']):.15e}")


if __name__ == "__main__":
    for name, (p, eps) in cases():
        dump(name, p, eps)
