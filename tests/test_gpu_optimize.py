"""north_star's second parity criterion: the optimized pulses after a fixed iteration count agree
to 1e-6 between the CUDA engine and the CPU oracle when both are driven by the same host loop
(grape.jl_b200/optimize.py: transcription of ext/GRAPELBFGSBExt.jl:18-147 on L-BFGS-B 3.0 `setulb`,
including the reference's nbd=3 / u=+Inf encoding).  Also the reference's own RNG-free end-to-end
pins, run on the GPU engine."""
import numpy as np
import pytest

from grape.jl_b200 import configs
from grape.jl_b200.optimize import (optimize, hamiltonian, Trajectory, J_T_sm, J_T_ss, J_a_fluence, QuadraticForm,
                                    ShapedAmplitude, Control)
from tests.oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu
PULSE_TOL = 1e-6
ITERS = 5


def _tls():
    eps = lambda t: 0.2 * float(configs.flattop(np.array([t]), T=5.0, t_rise=0.3)[0])
    H = hamiltonian(-0.5 * np.diag([1.0, -1.0]), ([[0, 1], [1, 0]], eps))
    return [Trajectory([1, 0], H, target_state=[0, 1])], np.linspace(0, 5, 501)


def _both(make, **kw):
    trajs, tlist = make()
    gpu = optimize(trajs, tlist, iter_stop=ITERS, **kw)
    trajs, tlist = make()
    cpu = optimize(trajs, tlist, iter_stop=ITERS, engine_factory=OracleEngine, **kw)
    assert gpu.iter == cpu.iter == ITERS
    for a, b in zip(gpu.optimized_controls, cpu.optimized_controls):
        assert np.max(np.abs(a - b)) <= PULSE_TOL * max(1.0, np.max(np.abs(b)))
    assert abs(gpu.J_T - cpu.J_T) <= 1e-8
    return gpu, cpu


def test_tls_pulses_after_five_iterations(lib_built):
    gpu, _ = _both(_tls, J_T=J_T_sm)
    # test/test_tls_optimization.jl:169-170
    assert gpu.J_T < 1e-3 and 0.75 < np.max(np.abs(gpu.optimized_controls[0])) < 0.85


def test_tls_bounded_and_taylor(lib_built):
    gpu, _ = _both(_tls, J_T=J_T_sm, upper_bound=0.7, lower_bound=-0.7)
    assert 0.65 < np.max(np.abs(gpu.optimized_controls[0])) < 0.700001     # test_tls_optimization.jl:259-260
    gt, _ = _both(_tls, J_T=J_T_sm, gradient_method="taylor")
    g0, _ = _both(_tls, J_T=J_T_sm)
    assert abs(gt.J_T - g0.J_T) < 1e-10                                     # test_tls_optimization.jl:229


def test_readme_example_converges_on_gpu(lib_built):
    H = hamiltonian([[1, 0], [0, -1]], ([[0, 1], [1, 0]], lambda t: 0.2))
    res = optimize([Trajectory([1, 0], H, target_state=[0, 1])], np.linspace(0, 5, 501), J_T=J_T_sm,
                   prop_method="B200ExpProp", check_convergence=lambda r: (r.J_T < 1e-3) and "J_T < 10⁻³")
    assert res.converged and res.J_T < 1e-3 and res.message == "J_T < 10⁻³"   # test_readme_example.jl:37-38


def _transmon():
    p, eps = configs.c2_transmon(NT=400)
    NT = p.NT
    cx, cy = Control(eps[:NT]), Control(eps[NT:])
    H = hamiltonian(p.H0[0], (p.Hc[0, 0], cx), (p.Hc[0, 1], cy))
    return [Trajectory(p.psi0[k], H, target_state=p.tgt[k]) for k in range(p.K)], p.tlist


def test_transmon_gate_pulses(lib_built):
    _both(_transmon, J_T=J_T_sm)


def _lambda_ensemble():
    p, eps = configs.c3_ensemble(n_delta=3, n_amp=3, NT=200)
    NT = p.NT
    cP, cS = Control(eps[:NT]), Control(eps[NT:])
    trajs = []
    for k in range(p.K):
        H = hamiltonian(p.H0[k], (p.Hc[k, 0], cP), (p.Hc[k, 1], cS))
        trajs.append(Trajectory(p.psi0[k], H, target_state=p.tgt[k]))
    return trajs, p.tlist


def test_ensemble_with_running_costs(lib_built):
    D = np.diag([0.0, 1.0, 0.0])
    _both(_lambda_ensemble, J_T=J_T_ss, J_a=J_a_fluence, lambda_a=0.01, g_b=QuadraticForm(D), lambda_b=0.2)


def test_cnot_saddle_point_on_gpu(lib_built):
    """The reference's RNG-free CNOT pin (test/test_lbfgsb_saddle_point.jl:89-124) through the CUDA engine (K=4, N=4,
    L=6, NT=1000): stalls at the J_T = 0.75 saddle with loose L-BFGS-B tolerances, passes it with the defaults; and
    the GPU-driven run follows the oracle-driven run (same iteration count, J_T to 1e-8) while it sits on the saddle."""
    from tests.oracle_engine import COracleEngine
    from tests.saddle_fixture import cnot_trajectories
    tr, tl = cnot_trajectories()
    g = optimize(tr, tl, J_T=J_T_sm, iter_stop=50, lbfgsb_pgtol=1e-5, lbfgsb_factr=1e7)
    assert not g.converged
    assert "NORM OF PROJECTED GRADIENT <= PGTOL" in g.message.replace("_", " ")
    assert abs(g.J_T - 0.75) < 1e-3
    tr, tl = cnot_trajectories()
    c = optimize(tr, tl, J_T=J_T_sm, iter_stop=50, lbfgsb_pgtol=1e-5, lbfgsb_factr=1e7, engine_factory=COracleEngine)
    assert g.iter == c.iter and abs(g.J_T - c.J_T) < 1e-8
    for a, b in zip(g.optimized_controls, c.optimized_controls):
        assert np.max(np.abs(a - b)) <= PULSE_TOL * max(1.0, np.max(np.abs(b)))
    tr, tl = cnot_trajectories()
    g = optimize(tr, tl, J_T=J_T_sm, iter_stop=50)
    assert g.converged and g.J_T < 1e-2
