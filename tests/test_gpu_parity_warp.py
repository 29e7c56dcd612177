"""GPU parity: sub-warp path (5 <= N <= 32) vs the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import grape_oracle as go
from tests.test_gpu_parity_small import check, engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [5, 6, 8, 9, 12, 16, 17, 24, 32])
def test_random_warp_sizes(lib_built, N):
    p, eps = configs.random_problem(K=3, N=N, L=2, NT=7, seed=300 + N, functional=gb.SM)
    p.tlist = p.tlist * (2.0 / np.sqrt(N))       # keep ||H dt|| moderate
    check(p, eps)


@pytest.mark.parametrize("functional", [gb.SM, gb.RE, gb.SS])
def test_warp_functionals_weights_shared(lib_built, functional):
    w = np.linspace(0.5, 1.5, 5)
    p, eps = configs.random_problem(K=5, N=6, L=3, NT=11, G=2, seed=40 + functional, hermitian=False,
                                    shaped=True, weights=w, functional=functional)
    p.tlist = p.tlist * 0.5
    check(p, eps)


@pytest.mark.parametrize("L", [1, 2, 3, 5, 7])
def test_warp_control_counts(lib_built, L):
    p, eps = configs.random_problem(K=2, N=7, L=L, NT=6, seed=500 + L)
    p.tlist = p.tlist * 0.5
    check(p, eps)


def test_warp_forced_on_small_n_matches_small_path(lib_built):
    p, eps = configs.random_problem(K=6, N=3, L=2, NT=12, seed=77, path=gb.PATH_WARP)
    check(p, eps)
    p, eps = configs.random_problem(K=6, N=4, L=2, NT=12, seed=78, path=gb.PATH_WARP, functional=gb.SS)
    check(p, eps)


def test_warp_running_costs_and_taylor(lib_built):
    N = 6
    rng = np.random.default_rng(9)
    A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    D = (A + A.conj().T) / 4
    p, eps = configs.random_problem(K=4, N=N, L=2, NT=10, seed=60, functional=gb.SS,
                                    gb_kind=gb.GB_QUADFORM, gb_D=D, lambda_b=0.4,
                                    ja_kind=gb.JA_FLUENCE, lambda_a=0.3)
    p.tlist = p.tlist * 0.5
    check(p, eps)
    p, eps = configs.random_problem(K=3, N=N, L=2, NT=10, seed=61, gradient_method=gb.TAYLOR)
    p.tlist = p.tlist * 0.5
    check(p, eps)


def test_warp_large_norm(lib_built):
    p, eps = configs.random_problem(K=2, N=8, L=2, NT=5, seed=62, uniform=True)
    p.tlist = p.tlist * 10.0
    check(p, eps, rtol=1e-9)


def test_c2_reduced_and_readbacks(lib_built):
    p, eps = configs.c2_transmon(NT=200)
    e, ref = check(p, eps)
    G = np.zeros_like(eps)
    e.evaluate_gradient(G, eps)
    assert np.max(np.abs(e.final_states() - ref["final_states"])) < 1e-12
    assert np.max(np.abs(e.stored_states(2) - ref["storage"][2])) < 1e-12
    assert np.max(np.abs(e.tau_grads(1) - ref["tau_grads"][1])) < 1e-12
    chi, rho = e.chi_states()
    assert np.max(np.abs(chi - ref["chi_states"])) < 1e-12


def test_c2_full_size_properties(lib_built):
    """Full BASELINE size (K=4, N=6, NT=2000): unitarity of the stored states and
    gradient vs. central finite differences of the engine's own functional."""
    p, eps = configs.c2_transmon()
    e = engine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    for k in range(4):
        st = e.stored_states(k)
        assert np.max(np.abs(np.linalg.norm(st, axis=0) - 1.0)) < 1e-11
    for i in (0, 777, 1999, 2500, 3999):
        x = eps.copy(); x[i] += 1e-5
        Jp = e.evaluate_functional(x)
        x[i] -= 2e-5
        Jm = e.evaluate_functional(x)
        assert abs((Jp - Jm) / 2e-5 - G[i]) < 1e-8
