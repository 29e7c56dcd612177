"""GPU parity of the real-symmetric schedule of the small path (csrc/small_sym.cuh: segment
propagators as cos/sin polynomials of a real matrix, gradient contraction from Im M only)
against the CPU oracle, against the general Hermitian schedule (GRAPE_B200_SEG_REAL=0) and
across Taylor orders, segment lengths and lane mappings.  Tolerance 1e-10 relative (north_star)."""
import os

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import grape_oracle as go

pytestmark = pytest.mark.gpu
RTOL = 1e-10


class env:
    """temporarily set environment variables read at engine creation"""

    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}
        self.old = {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def engine(p, **kv):
    from grape.jl_b200.engine import GrapeEngine
    # small test ensembles do not reach the size at which the fused formation kernel is chosen: force it
    with env(GRAPE_B200_FORCE_FORMSEG=1, **kv):
        return GrapeEngine(p)


def check(p, eps, schedule=3, **kv):
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    e = engine(p, **kv)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    assert e.small_schedule() == schedule, e.small_schedule()
    scale = max(np.max(np.abs(ref["G"])), 1e-6)
    assert abs(J - ref["J"]) <= RTOL * max(1.0, abs(ref["J"]))
    assert np.max(np.abs(G - ref["G"])) <= RTOL * scale
    assert np.max(np.abs(e.tau_vals - ref["tau"])) <= RTOL
    for k in range(min(p.K, 3)):
        assert np.max(np.abs(e.stored_states(k) - ref["storage"][k])) <= 1e-12
        # the dump re-runs the contraction with the general Hermitian kernel
        assert np.max(np.abs(e.tau_grads(k) - ref["tau_grads"][k])) <= 1e-12 * max(1.0, scale)
    # ... and must not disturb the next evaluation
    G2 = np.zeros_like(eps)
    J2 = e.evaluate_gradient(G2, eps)
    assert J2 == J and np.array_equal(G, G2)
    assert e.small_schedule() == schedule
    return e, ref, G


@pytest.mark.parametrize("seg_len", [2, 7, 64])
@pytest.mark.parametrize("N", [1, 2, 3])
def test_real_symmetric_small_theta(lib_built, seg_len, N):
    p, eps = configs.random_problem(K=5, N=N, L=2, NT=37, seed=400 + N, real=True, shaped=True, functional=gb.SM)
    p.tlist[:] = p.tlist * 0.005     # theta ~ 1e-3 .. 1e-2: orders 5..8, ragged last segment
    check(p, eps, GRAPE_B200_SEG_S=seg_len)[0].close()


@pytest.mark.parametrize("scale,order_hint", [(1e-9, 2), (1e-6, 3), (3e-5, 4), (3e-4, 5), (1.5e-3, 6), (4e-3, 7), (9e-3, 8)])
def test_every_taylor_order_template(lib_built, scale, order_hint):
    """||H dt|| swept over seven decades: every sym_step<N, M>, M = 2..8, is executed"""
    p, eps = configs.random_problem(K=9, N=3, L=2, NT=29, seed=410 + order_hint, real=True, uniform=True,
                                    functional=gb.SS)
    p.tlist[:] = p.tlist * (scale / 0.05) / 4.0
    check(p, eps)[0].close()


@pytest.mark.parametrize("K", [1, 3, 31, 33, 70])
def test_lane_mapping_and_functionals(lib_built, K):
    for fn, L in ((gb.SM, 1), (gb.RE, 3), (gb.SS, 5)):
        p, eps = configs.random_problem(K=K, N=3, L=L, NT=21, seed=420 + K, real=True, functional=fn, G=min(K, 3),
                                        weights=np.linspace(0.5, 1.5, K))
        p.tlist[:] = p.tlist * 0.01
        check(p, eps)[0].close()


def test_ineligible_steps_fall_back_on_the_device(lib_built):
    """ROUND-1 kernels (GRAPE_B200_SYM_V=1): a step with ||H dt|| > 0.099 (economised polynomial; Taylor: 0.0308 -- more
    than 8 orders / sub-stepping) switches
    the whole call to the general Hermitian kernel on the device; the flag follows the pulses call by call"""
    p, eps = configs.random_problem(K=6, N=3, L=2, NT=25, seed=431, real=True, functional=gb.SM)
    check(p, eps, schedule=2, GRAPE_B200_SYM_V=1)[0].close()              # dt ~ 0.05, ||H|| ~ 3
    p.tlist[:] = p.tlist * 0.01
    e = engine(p, GRAPE_B200_SYM_V=1)
    op = go.from_problem(p)
    G = np.zeros_like(eps)
    for amp, sched in ((1.0, 3), (150.0, 2), (0.5, 3), (150.0, 2), (1.0, 3)):
        x = eps * amp
        ref = go.evaluate_gradient(op, x)
        J = e.evaluate_gradient(G, x)
        assert e.small_schedule() == sched
        assert abs(J - ref["J"]) <= RTOL
        assert np.max(np.abs(G - ref["G"])) <= RTOL * max(np.max(np.abs(ref["G"])), 1e-6)
    e.close()


@pytest.mark.parametrize("scan", [1, 0])
def test_large_steps_are_sub_stepped_inside_the_real_kernel(lib_built, scan):
    """Staged kernels (round 2): steps with ||H dt|| > 0.0308 are served by the real-symmetric gradient kernel itself,
    as equal sub-steps of 8 orders, per warp and per step -- no call-wide switch to the complex kernel.  ||H dt|| up
    to ~ 6 here (about 200 sub-steps), mixed with small steps in the same call."""
    p, eps = configs.random_problem(K=6, N=3, L=2, NT=25, seed=431, real=True, functional=gb.SM)
    check(p, eps, schedule=3, GRAPE_B200_SEG_SCAN=scan)[0].close()        # dt ~ 0.05, ||H|| ~ 3: every step sub-stepped
    p.tlist[:] = p.tlist * 0.01
    e = engine(p, GRAPE_B200_SEG_SCAN=scan)
    op = go.from_problem(p)
    G = np.zeros_like(eps)
    rng = np.random.default_rng(3)
    for amp in (1.0, 40.0, 0.5, 40.0, 1.0):
        x = eps * amp
        x[rng.integers(0, len(x), 7)] *= 300.0 / amp                       # a few very large steps among small ones
        ref = go.evaluate_gradient(op, x)
        J = e.evaluate_gradient(G, x)
        assert e.small_schedule() == 3
        assert abs(J - ref["J"]) <= RTOL
        assert np.max(np.abs(G - ref["G"])) <= RTOL * max(np.max(np.abs(ref["G"])), 1e-6)
    e.close()


@pytest.mark.parametrize("NT,S", [(37, 2), (200, 2), (129, 1), (1000, 8), (1000, 11), (333, 16), (64, 64)])
def test_scan_schedule_segment_counts(lib_built, NT, S):
    """prefix products by the warp scan (csrc/small_sym.cuh small_formscan_sym): 1..4 warps per generator block,
    ragged last segment, a forced segment length too short for one block (S = 1, 2 with NT > 128: widened)"""
    p, eps = configs.random_problem(K=7, N=3, L=2, NT=NT, seed=470 + NT % 7, real=True, functional=gb.SS, G=3,
                                    weights=np.linspace(0.5, 1.5, 7))
    p.tlist[:] = p.tlist * 0.004
    check(p, eps, GRAPE_B200_SEG_S=S)[0].close()


@pytest.mark.parametrize("N,L", [(1, 1), (2, 1), (2, 3), (3, 4)])
def test_scan_schedule_sizes_and_control_counts(lib_built, N, L):
    p, eps = configs.random_problem(K=9, N=N, L=L, NT=150, seed=480 + N + L, real=True, functional=gb.SM, shaped=True)
    p.tlist[:] = p.tlist * 0.004
    check(p, eps)[0].close()
    check(p, eps, GRAPE_B200_SEG_SCAN=0)[0].close()      # same kernels fed by the boundary chains


def test_complex_hermitian_and_taylor_do_not_take_the_real_path(lib_built):
    p, eps = configs.random_problem(K=5, N=3, L=2, NT=19, seed=441, hermitian=True, functional=gb.SM)
    p.tlist[:] = p.tlist * 0.01
    check(p, eps, schedule=2)[0].close()
    p, eps = configs.random_problem(K=5, N=3, L=2, NT=19, seed=442, real=True, gradient_method=gb.TAYLOR)
    p.tlist[:] = p.tlist * 0.01
    check(p, eps, schedule=2)[0].close()               # boundaries still come from the real-symmetric scan kernels
    check(p, eps, schedule=2, GRAPE_B200_SEG_SCAN=0)[0].close()
    p, eps = configs.random_problem(K=5, N=3, L=2, NT=19, seed=443, real=True)
    p.tlist[:] = p.tlist * 0.01
    check(p, eps, schedule=2, GRAPE_B200_SEG_REAL=0)[0].close()


def test_both_register_allocations(lib_built):
    p, eps = configs.c3_ensemble(n_delta=6, n_amp=7, NT=400)
    for occ in (2, 3):
        check(p, eps, GRAPE_B200_SYM_OCC=occ)[0].close()


def test_host_chi_and_functional_only(lib_built):
    p, eps = configs.c3_ensemble(n_delta=4, n_amp=4, NT=400, functional=gb.HOST)   # ||H dt|| <= 0.023
    pr, _ = configs.c3_ensemble(n_delta=4, n_amp=4, NT=400, functional=gb.SS)
    ref = go.evaluate_gradient(go.from_problem(pr), eps)
    e = engine(p)
    e.forward(eps)
    tau = np.einsum("ki,ki->k", p.tgt.conj(), e.final_states())
    Gp = np.zeros_like(eps)
    e.backward_chi((tau / p.K)[:, None] * p.tgt, Gp)
    assert e.small_schedule() == 3
    assert np.max(np.abs(Gp - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    e.close()
    e = engine(pr)
    assert abs(e.evaluate_functional(eps) - ref["J"]) <= RTOL
    assert np.max(np.abs(e.stored_states(5) - ref["storage"][5])) <= 1e-12
    e.close()


def test_c3_full_size_real_vs_hermitian_schedule(lib_built):
    """BASELINE configs[2] at full size (fused formation chosen by the size rule, no override): the real-symmetric
    schedule serves the call and agrees with the general Hermitian schedule to 1e-12 on every gradient element"""
    from grape.jl_b200.engine import GrapeEngine
    p, eps = configs.c3_ensemble()
    e = GrapeEngine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    assert e.small_schedule() == 3
    e.close()
    with env(GRAPE_B200_SEG_REAL=0):
        e0 = GrapeEngine(p)
    G0 = np.zeros_like(eps)
    J0 = e0.evaluate_gradient(G0, eps)
    assert e0.small_schedule() == 2
    e0.close()
    assert abs(J - J0) < 1e-13
    assert np.max(np.abs(G - G0)) <= 1e-12 * np.max(np.abs(G0))
    sub, _ = configs.c3_ensemble(n_delta=2, n_amp=64)           # first 128 trajectories against the oracle
    ref = go.evaluate_gradient(go.from_problem(sub), eps)
    es = engine(sub)
    Gs = np.zeros_like(eps)
    es.evaluate_gradient(Gs, eps)
    assert np.max(np.abs(Gs - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    es.close()


@pytest.mark.parametrize("econ", [1, 0])
@pytest.mark.parametrize("theta", [3e-4, 1e-3, 1e-2, 0.05, 0.3, 0.8, 2.5])
def test_propagator_classes_economised_and_taylor(lib_built, theta, econ):
    """||H dt|| swept over the five polynomial classes of the formation kernels (degree 4, 6, 8, 12, 16) and into scaling
    and squaring, with the Chebyshev-cut coefficients (default) and the Taylor coefficients (GRAPE_B200_ECON=0): both
    match the oracle's exact exponential at 1e-10 (csrc/small_sym.cuh sym_tables_upload, csrc/econ.cuh)."""
    for scan in (1, 0):
        p, eps = configs.random_problem(K=7, N=3, L=2, NT=23, seed=470, real=True, uniform=True, functional=gb.SM)
        nrm = max(np.abs(p.H0[g] + sum(abs(eps.reshape(p.L, p.NT)[l]).max() * np.abs(p.Hc[g, l]) for l in range(p.L))).sum(0).max()
                  for g in range(p.H0.shape[0]))
        p.tlist[:] = p.tlist * (theta / (nrm * 0.05))
        check(p, eps, GRAPE_B200_ECON=econ, GRAPE_B200_SEG_SCAN=scan)[0].close()


def test_economised_and_taylor_series_agree_on_c3_shard(lib_built):
    """a 256-trajectory shard of C3, all 1000 steps: the economised polynomial (degree 6 per gradient step) and the
    Taylor series (degree 7) give the same J and gradient to 1e-12"""
    p, eps = configs.c3_ensemble(n_delta=16, n_amp=16)
    out = {}
    for econ in (1, 0):
        e = engine(p, GRAPE_B200_ECON=econ)
        G = np.zeros_like(eps)
        out[econ] = (e.evaluate_gradient(G, eps), G)
        assert e.small_schedule() == 3
        e.close()
    scale = np.max(np.abs(out[0][1]))
    assert abs(out[1][0] - out[0][0]) <= 1e-12 * max(1.0, abs(out[0][0]))
    assert np.max(np.abs(out[1][1] - out[0][1])) <= 1e-12 * scale
