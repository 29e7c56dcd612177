"""CPU: the spectral dense oracle (oracle/dense_oracle.py), which produces the full-size C4 / C5
golden vectors, agrees with the literal restatement of the reference (oracle/grape_oracle.py:
Pade expm of the N x N and N(L+1) x N(L+1) matrices per trajectory-step) on reduced sizes, for
both gradient methods, every built-in functional and both running costs."""
import os

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import dense_oracle as do
from oracle import grape_oracle as go

HERE = os.path.dirname(os.path.abspath(__file__))


def _cmp(r, d, rtol=1e-12):
    assert abs(r["J"] - d["J"]) <= rtol
    assert np.max(np.abs(r["J_parts"] - d["J_parts"])) <= rtol
    assert np.max(np.abs(r["tau"] - d["tau"])) <= rtol
    sc = np.max(np.abs(r["G"]))
    assert np.max(np.abs(r["G"] - d["G"])) <= rtol * sc
    assert np.max(np.abs(r["grad_J_Tb"] - d["grad_J_Tb"])) <= rtol * sc
    assert np.max(np.abs(r["final_states"] - d["final_states"])) <= rtol
    assert np.max(np.abs(r["chi_norms"] - d["chi_norms"])) <= rtol


@pytest.mark.parametrize("functional", [gb.SM, gb.RE, gb.SS])
def test_spectral_oracle_matches_literal_oracle_c4_shape(functional):
    p, eps = configs.c4_dense450(N=24, K=5, NT=9, functional=functional)
    _cmp(go.evaluate_gradient(go.from_problem(p), eps), do.evaluate_gradient(p, eps))


@pytest.mark.parametrize("method", [gb.GRADGEN, gb.TAYLOR])
def test_spectral_oracle_matches_literal_oracle_c5_costs(method):
    p, eps = configs.c5_dense1024(N=20, K=6, NT=11, gradient_method=method)
    assert p.gb_kind == gb.GB_QUADFORM and p.ja_kind == gb.JA_FLUENCE
    _cmp(go.evaluate_gradient(go.from_problem(p), eps), do.evaluate_gradient(p, eps))


def test_spectral_oracle_nonuniform_grid_shapes_weights_degenerate_levels():
    rng = np.random.default_rng(5)
    p, eps = configs.random_problem(K=4, N=6, L=3, NT=13, G=1, seed=91, shaped=True, functional=gb.SS,
                                    weights=np.linspace(0.5, 1.5, 4))
    # exactly degenerate drift levels: the divided differences must not cancel
    p.H0[0] = np.diag([0.3, 0.3, 0.3, -1.0, 2.0, 2.0]).astype(complex)
    _cmp(go.evaluate_gradient(go.from_problem(p), eps), do.evaluate_gradient(p, eps, keep_eig=False))


@pytest.mark.parametrize("name", ["c4_dense450_full", "c5_dense1024_full"])
def test_full_size_dense_golden_is_committed_and_consistent(name):
    """The full-size fixtures store G, J_parts, tau and a fingerprint of the (seeded, regenerated) operators."""
    z = np.load(os.path.join(HERE, "golden", "dense_full", name + ".npz"))
    p, eps = (configs.c4_dense450 if name.startswith("c4") else configs.c5_dense1024)()
    assert z["G"].shape == (p.L * p.NT,) and z["tau"].shape == (p.K,)
    assert np.array_equal(z["pulsevals"], eps)
    fp = np.array([np.sum(p.H0).real, np.sum(np.abs(p.Hc)), np.sum(p.tgt).imag])
    assert np.allclose(z["fingerprint"], fp, rtol=0, atol=1e-9)
    assert abs(float(z["J"]) - float(np.sum(z["J_parts"]))) < 1e-15
    assert np.all(np.isfinite(z["G"])) and np.max(np.abs(z["G"])) > 0
