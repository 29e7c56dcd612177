"""Per-term amplitudes and non-linear controls (reference: the amplitude belongs to each term of each generator;
`get_control_derivs`, src/workspace.jl:283-285; mu evaluated per step and the `isnothing(mu)` -> 0 branch,
src/optimize.jl:946-951).  CPU: host chain rule + oracle against finite differences of J; GPU: the CUDA engine's
amplitude mode (grape_b200_eval_fg_amplitudes) against the oracle at 1e-10."""
import numpy as np
import pytest

from grape.jl_b200.optimize import (GrapeWrk, hamiltonian, Trajectory, Control, ShapedAmplitude, NonlinearAmplitude,
                                    J_T_sm, J_T_ss, J_a_fluence, optimize)
from tests.oracle_engine import OracleEngine


def _ops(N, rng, n):
    out = []
    for _ in range(n):
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        out.append((A + A.conj().T) / 2)
    return out


def _mixed_problem(N=3, NT=24, seed=1, single=False):
    """two controls; generator A: c0 linear + c0 squared (non-linear) + c1 shaped; generator B: c0 shaped differently,
    no dependence on c1 at all (mu = nothing -> zero gradient contribution); generator C: c1 unshaped."""
    rng = np.random.default_rng(seed)
    tlist = np.concatenate([[0.0], np.cumsum(0.02 + 0.03 * rng.random(NT))])
    c0 = Control(lambda t: 0.4 * np.sin(3 * t) + 0.2)
    c1 = Control(lambda t: 0.3 * np.cos(2 * t))
    H0a, H0b, H0c, A1, A2, A3, B1, C1 = _ops(N, rng, 8)
    sq = NonlinearAmplitude(c0, lambda e, t: e * e, lambda e, t: 2.0 * e)
    shp1 = ShapedAmplitude(c1, lambda t: 0.5 + 0.5 * np.sin(t) ** 2)
    shp0 = ShapedAmplitude(c0, lambda t: np.exp(-t))
    gA = hamiltonian(H0a, (A1, c0), (A2, sq), (A3, shp1))
    gB = hamiltonian(H0b, (B1, shp0))
    gC = hamiltonian(H0c, (C1, c1))

    def st():
        v = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        return v / np.linalg.norm(v)
    if single:      # one shared generator (the dense path's requirement)
        gB = gC = gA
    trajs = [Trajectory(st(), gA, target_state=st()), Trajectory(st(), gB, target_state=st(), weight=0.7),
             Trajectory(st(), gC, target_state=st()), Trajectory(st(), gA, target_state=st(), weight=1.3)]
    return trajs, tlist


def _fd(wrk, x, idx, h=1e-6):
    out = []
    for i in idx:
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        out.append((wrk.evaluate_functional(xp, False) - wrk.evaluate_functional(xm, False)) / (2 * h))
    return np.array(out)


@pytest.mark.parametrize("JT", [J_T_sm, J_T_ss])
def test_amplitude_mode_gradient_matches_finite_differences(JT):
    trajs, tlist = _mixed_problem()
    wrk = GrapeWrk(trajs, tlist, dict(J_T=JT, J_a=J_a_fluence, lambda_a=0.3), engine_factory=OracleEngine)
    assert wrk.amplitude_mode and len(wrk.slots.slots) == 5 and wrk.problem.L == 5 and wrk.problem.ja_kind == 0
    x = wrk.pulsevals.copy()
    G = np.zeros_like(x)
    J = wrk.evaluate_gradient(G, x)
    assert abs(J - wrk.evaluate_functional(x, False)) < 1e-13
    idx = [0, 5, 11, 23, 24, 30, 47]
    assert np.max(np.abs(_fd(wrk, x, idx) - G[idx])) < 2e-8 * max(1.0, np.max(np.abs(G)))
    assert wrk.J_parts[1] == pytest.approx(0.3 * J_a_fluence(x, tlist))


def test_same_control_shaped_in_one_generator_and_unshaped_in_another():
    """the shape belongs to the amplitude (term), not to the control.  One shaped and one unshaped trajectory
    sharing a control must see different amplitudes: J depends on which one is shaped."""
    rng = np.random.default_rng(3)
    N, NT = 3, 20
    tlist = np.linspace(0, 1.0, NT + 1)
    c = Control(lambda t: 0.5)
    H0, H1 = _ops(N, rng, 2)
    shaped = hamiltonian(H0, (H1, ShapedAmplitude(c, lambda t: 0.2 + t)))
    plain = hamiltonian(H0, (H1, c))
    a, b = np.eye(N)[0].astype(complex), np.eye(N)[1].astype(complex)
    t1, t2 = np.eye(N)[2].astype(complex), (np.eye(N)[0] + np.eye(N)[1]) / np.sqrt(2)
    J = {}
    for name, (g1, g2) in dict(first_shaped=(shaped, plain), second_shaped=(plain, shaped)).items():
        wrk = GrapeWrk([Trajectory(a, g1, target_state=t1), Trajectory(b, g2, target_state=t2)], tlist,
                       dict(J_T=J_T_ss), engine_factory=OracleEngine)
        assert wrk.amplitude_mode and len(wrk.slots.slots) == 2
        x = wrk.pulsevals.copy()
        G = np.zeros_like(x)
        J[name] = wrk.evaluate_gradient(G, x)
        assert np.max(np.abs(_fd(wrk, x, [0, 7, 19]) - G[[0, 7, 19]])) < 2e-8
    assert abs(J["first_shaped"] - J["second_shaped"]) > 1e-3
    # consistently shaped controls stay on the direct path (descriptor shape, device J_a)
    wrk = GrapeWrk([Trajectory(a, shaped, target_state=t1), Trajectory(b, shaped, target_state=t2)], tlist,
                   dict(J_T=J_T_ss, J_a=J_a_fluence), engine_factory=OracleEngine)
    assert not wrk.amplitude_mode and wrk.problem.shape is not None and wrk.problem.ja_kind == 1


@pytest.mark.gpu
@pytest.mark.parametrize("N", [2, 3, 6, 40])
def test_cuda_amplitude_mode_matches_oracle(lib_built, N):
    """small (real-symmetric / complex), sub-warp and dense paths"""
    trajs, tlist = _mixed_problem(N=N, NT=16 if N == 40 else 40, seed=N, single=N > 32)
    res = {}
    for name, fac in (("gpu", None), ("cpu", OracleEngine)):
        wrk = GrapeWrk(trajs, tlist, dict(J_T=J_T_sm, J_a=J_a_fluence, lambda_a=0.3), engine_factory=fac)
        x = wrk.pulsevals.copy()
        G = np.zeros_like(x)
        J = wrk.evaluate_gradient(G, x)
        res[name] = (J, G.copy(), wrk.evaluate_functional(x, False), wrk.result.tau_vals.copy())
    (Jg, Gg, Fg, tg), (Jc, Gc, Fc, tc) = res["gpu"], res["cpu"]
    assert abs(Jg - Jc) <= 1e-10 and abs(Fg - Fc) <= 1e-10 and np.max(np.abs(tg - tc)) <= 1e-10
    assert np.max(np.abs(Gg - Gc)) <= 1e-10 * np.max(np.abs(Gc))


@pytest.mark.gpu
def test_cuda_optimization_with_nonlinear_control(lib_built):
    """five L-BFGS-B iterations with a squared-amplitude drive: GPU-driven and oracle-driven pulses agree to 1e-6"""
    def make():
        c = Control(lambda t: 0.6 + 0.2 * np.sin(t))
        sq = NonlinearAmplitude(c, lambda e, t: e * e, lambda e, t: 2.0 * e)
        H = hamiltonian(-0.5 * np.diag([1.0, -1.0]), ([[0, 1], [1, 0]], sq))
        return [Trajectory([1, 0], H, target_state=[0, 1])], np.linspace(0, 5, 201)
    tr, tl = make()
    g = optimize(tr, tl, J_T=J_T_sm, iter_stop=5)
    tr, tl = make()
    c = optimize(tr, tl, J_T=J_T_sm, iter_stop=5, engine_factory=OracleEngine)
    assert g.iter == c.iter == 5 and abs(g.J_T - c.J_T) <= 1e-8 and g.J_T < 0.9 * 1.0
    assert np.max(np.abs(g.optimized_controls[0] - c.optimized_controls[0])) <= 1e-6
