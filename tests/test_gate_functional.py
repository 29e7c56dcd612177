"""Gate functionals via U_L (reference docs/src/background.md:552-610; QuantumControl.Functionals.gate_functional /
make_gate_chi): J_T is a function of (U_L)_ij = <phi_i|Psi_j(T)>, chi_k(T) = -1/2 sum_i (nabla_U J_T)_ik |phi_i>,
served by the host-chi path (forward on the device, J_T / chi on the host, backward on the device)."""
import numpy as np
import pytest

from grape.jl_b200 import configs
from grape.jl_b200.optimize import (GrapeWrk, hamiltonian, Trajectory, Control, J_T_sm, gate_functional, make_gate_chi,
                                    logical_gate, optimize)
from tests.oracle_engine import OracleEngine


def _transmon(NT=60):
    p, eps = configs.c2_transmon(NT=NT)
    cx, cy = Control(eps[:NT]), Control(eps[NT:] + 0.02)
    H = hamiltonian(p.H0[0], (p.Hc[0, 0], cx), (p.Hc[0, 1], cy))
    basis = [np.eye(p.N, dtype=complex)[i] for i in range(2)]            # logical subspace: the two lowest levels
    O = np.array([[0, 1], [1, 0]], dtype=complex)                         # target gate X
    trajs = [Trajectory(basis[k], H, target_state=sum(O[i, k] * basis[i] for i in range(2))) for k in range(2)]
    return trajs, p.tlist, basis, O


def _sm(O):
    d = O.shape[0]
    return (lambda U: 1.0 - abs(np.vdot(O, U)) ** 2 / d ** 2,            # J_T_U = 1 - |Tr(O^dagger U)|^2 / d^2
            lambda U: -2.0 * np.vdot(O, U) * O / d ** 2)                  # nabla_U J_T = 2 dJ/dU* = -2 Tr(O^dagger U) O / d^2


def _pe(O):
    """a functional that is NOT expressible through the overlaps tau_k alone: population-weighted gate error
    J = 1 - sum_ij w_ij |U_ij|^2 / d with w = |O|^2 + 0.3 (off-target leakage inside the subspace is penalised less)"""
    W = np.abs(O) ** 2 + 0.3
    d = O.shape[0]
    return (lambda U: 1.0 - float(np.sum(W * np.abs(U) ** 2)) / d, lambda U: -2.0 * W * U / d)


def _fd(wrk, x, idx, h=1e-6):
    out = []
    for i in idx:
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        out.append((wrk.evaluate_functional(xp, False) - wrk.evaluate_functional(xm, False)) / (2 * h))
    return np.array(out)


def test_square_modulus_gate_functional_equals_J_T_sm():
    trajs, tlist, basis, O = _transmon()
    JU, gU = _sm(O)
    wg = GrapeWrk(trajs, tlist, dict(J_T=gate_functional(JU), chi=make_gate_chi(gU)), engine_factory=OracleEngine)
    ws = GrapeWrk(trajs, tlist, dict(J_T=J_T_sm), engine_factory=OracleEngine)
    x = wg.pulsevals.copy()
    Gg, Gs = np.zeros_like(x), np.zeros_like(x)
    assert abs(wg.evaluate_gradient(Gg, x) - ws.evaluate_gradient(Gs, x)) < 1e-13
    assert np.max(np.abs(Gg - Gs)) < 1e-13 * max(1.0, np.max(np.abs(Gs)))


def test_general_gate_functional_gradient_matches_finite_differences():
    trajs, tlist, basis, O = _transmon()
    JU, gU = _pe(O)
    w = GrapeWrk(trajs, tlist, dict(J_T=gate_functional(JU), chi=make_gate_chi(gU)), engine_factory=OracleEngine)
    x = w.pulsevals.copy()
    G = np.zeros_like(x)
    J = w.evaluate_gradient(G, x)
    U = logical_gate(w.engine.final_states(), basis)
    assert abs(J - JU(U)) < 1e-14 and U.shape == (2, 2)
    idx = [0, 17, 59, 60, 100]
    assert np.max(np.abs(_fd(w, x, idx) - G[idx])) < 2e-8


@pytest.mark.gpu
def test_cuda_gate_functional_matches_oracle_and_optimizes(lib_built):
    trajs, tlist, basis, O = _transmon(NT=200)
    JU, gU = _pe(O)
    kw = dict(J_T=gate_functional(JU), chi=make_gate_chi(gU))
    wg = GrapeWrk(trajs, tlist, kw)
    wc = GrapeWrk(trajs, tlist, kw, engine_factory=OracleEngine)
    x = wg.pulsevals.copy()
    Gg, Gc = np.zeros_like(x), np.zeros_like(x)
    assert abs(wg.evaluate_gradient(Gg, x) - wc.evaluate_gradient(Gc, x)) <= 1e-10
    assert np.max(np.abs(Gg - Gc)) <= 1e-10 * np.max(np.abs(Gc))
    g = optimize(trajs, tlist, iter_stop=5, **kw)
    c = optimize(trajs, tlist, iter_stop=5, engine_factory=OracleEngine, **kw)
    assert g.iter == c.iter == 5 and g.J_T < wg.J_parts[0] and abs(g.J_T - c.J_T) <= 1e-8
    for a, b in zip(g.optimized_controls, c.optimized_controls):
        assert np.max(np.abs(a - b)) <= 1e-6 * max(1.0, np.max(np.abs(b)))
