"""GPU parity: dense DMMA path (N > 32, or forced) vs the CPU oracle, through the C-ABI.
The case matrix over both backward forms lives in tests/test_gpu_parity_dense_krylov.py."""
import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import grape_oracle as go
from tests.test_gpu_parity_small import check, engine

pytestmark = pytest.mark.gpu


def test_dense_forced_on_small_problem(lib_built):
    p, eps = configs.random_problem(K=4, N=6, L=2, NT=8, G=1, seed=81, path=gb.PATH_DENSE)
    p.tlist = p.tlist * 0.5
    check(p, eps)


def test_dense_readbacks_and_split(lib_built):
    p, eps = configs.c4_dense450(N=50, K=7, NT=5)
    e, ref = check(p, eps)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    assert np.max(np.abs(e.final_states() - ref["final_states"])) < 1e-12
    assert np.max(np.abs(e.stored_states(3) - ref["storage"][3])) < 1e-12
    chi, rho = e.chi_states()
    assert np.max(np.abs(chi - ref["chi_states"])) < 1e-12
    sums = e.forward(eps)
    Gp = np.zeros_like(eps)
    e.backward(sums, Gp)
    # eval_fg runs the two chains concurrently (chi_k(T) applied in the contraction), the split call one after the
    # other: same series, different rounding
    assert np.max(np.abs(Gp - G)) <= 1e-12 * np.max(np.abs(G))


def test_c5_reduced_steps_full_width(lib_built):
    """C5 at full width (N=1024, K=64, J_a + g_b) on 2 time steps: oracle parity on a
    subset of the ensemble is too slow on CPU (N(L+1)=3072 expm), so compare against
    the oracle's :taylor variant which needs only N x N exponentials."""
    p, eps = configs.c5_dense1024(NT=2, K=8)
    ref = go.evaluate_gradient(go.from_problem(p, gradient_method=go.TAYLOR), eps)
    e = engine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    assert abs(J - ref["J"]) < 1e-10
    assert np.max(np.abs(G - ref["G"])) <= 1e-10 * np.max(np.abs(ref["G"]))


# ---- tiled DMMA kernels (csrc/dense2.cuh), forced on problems the oracle finishes quickly
@pytest.fixture
def dense2_forced():
    import os
    old = os.environ.get("GRAPE_B200_DENSE2")
    os.environ["GRAPE_B200_DENSE2"] = "1"
    yield
    if old is None:
        del os.environ["GRAPE_B200_DENSE2"]
    else:
        os.environ["GRAPE_B200_DENSE2"] = old


def test_dense2_substeps_and_running_costs(lib_built, dense2_forced):
    p, eps = configs.c4_dense450(N=48, K=10, NT=4)
    p.tlist = p.tlist * 8.0
    check(p, eps, rtol=1e-9)
    p, eps = configs.c5_dense1024(N=40, K=16, NT=7)
    e, ref = check(p, eps)
    assert e.J_parts[1] > 0 and e.J_parts[2] > 0
    assert np.max(np.abs(e.final_states() - ref["final_states"])) < 1e-12
    assert np.max(np.abs(e.stored_states(3) - ref["storage"][3])) < 1e-12


def test_dense2_matches_strip_kernels_c5_width(lib_built):
    """C5 at full width (N=1024, K=64, J_a + g_b), 2 time steps: tiled kernels (default there)
    against the strip kernels (GRAPE_B200_DENSE2=0)."""
    import os
    p, eps = configs.c5_dense1024(NT=2)
    e = engine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    e.close()
    os.environ["GRAPE_B200_DENSE2"] = "0"
    try:
        e0 = engine(p)
        G0 = np.zeros_like(eps)
        J0 = e0.evaluate_gradient(G0, eps)
        e0.close()
    finally:
        del os.environ["GRAPE_B200_DENSE2"]
    assert abs(J - J0) <= 1e-12
    assert np.max(np.abs(G - G0)) <= 1e-11 * np.max(np.abs(G0))
