"""GPU parity of the time-segmented small-N schedule (csrc/small_seg.cuh) against the CPU
oracle, against the plain-chain schedule (PATH_SMALL_CHAIN) and across segment lengths.
Tolerance 1e-10 relative (north_star)."""
import os

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import grape_oracle as go

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def engine(p):
    from grape.jl_b200.engine import GrapeEngine
    return GrapeEngine(p)


def run(p, eps):
    e = engine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    return e, J, G


def check(p, eps):
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    e, J, G = run(p, eps)
    scale = max(np.max(np.abs(ref["G"])), 1e-6)
    assert abs(J - ref["J"]) <= RTOL * max(1.0, abs(ref["J"]))
    assert np.max(np.abs(G - ref["G"])) <= RTOL * scale
    assert np.max(np.abs(e.tau_vals - ref["tau"])) <= RTOL
    for k in range(min(p.K, 3)):
        assert np.max(np.abs(e.stored_states(k) - ref["storage"][k])) <= 1e-12
        assert np.max(np.abs(e.tau_grads(k) - ref["tau_grads"][k])) <= 1e-12 * max(1.0, scale)
    return e, ref, G


@pytest.fixture
def seg_len(request):
    old = os.environ.get("GRAPE_B200_SEG_S")
    os.environ["GRAPE_B200_SEG_S"] = str(request.param)
    yield request.param
    if old is None:
        del os.environ["GRAPE_B200_SEG_S"]
    else:
        os.environ["GRAPE_B200_SEG_S"] = old


@pytest.mark.parametrize("seg_len", [2, 3, 7, 16, 64], indirect=True)
@pytest.mark.parametrize("N", [1, 2, 3, 4])
def test_segment_lengths_small_theta(lib_built, seg_len, N):
    """small ||H dt|| -> Krylov contraction (fast branch) for N <= 3; ragged last segment."""
    p, eps = configs.random_problem(K=5, N=N, L=2, NT=37, seed=100 + N, hermitian=False, shaped=True,
                                    functional=gb.SS)
    p.tlist[:] = p.tlist * 0.02     # theta ~ 1e-2
    check(p, eps)


@pytest.mark.parametrize("seg_len", [2, 5, 64], indirect=True)
def test_segment_lengths_large_theta(lib_built, seg_len):
    """||H dt|| ~ O(1): block-recursion branch, with sub-stepping."""
    p, eps = configs.random_problem(K=7, N=3, L=3, NT=23, seed=7, functional=gb.SM)
    check(p, eps)
    p, eps = configs.random_problem(K=3, N=2, L=1, NT=9, seed=8, functional=gb.RE)
    p.tlist[:] = p.tlist * 30.0
    check(p, eps)


@pytest.mark.parametrize("K", [1, 2, 3, 31, 32, 33, 70])
def test_trajectory_counts_lane_mapping(lib_built, K):
    p, eps = configs.random_problem(K=K, N=3, L=2, NT=21, seed=K, functional=gb.SM, G=min(K, 3))
    p.tlist[:] = p.tlist * 0.05
    check(p, eps)


def test_matches_plain_chain_schedule(lib_built):
    for fn in (gb.SM, gb.RE, gb.SS):
        p, eps = configs.c3_ensemble(n_delta=5, n_amp=9, NT=130, functional=fn)
        _, J1, G1 = run(p, eps)
        p2, _ = configs.c3_ensemble(n_delta=5, n_amp=9, NT=130, functional=fn, path=gb.PATH_SMALL_CHAIN)
        _, J2, G2 = run(p2, eps)
        assert abs(J1 - J2) <= 1e-13
        assert np.max(np.abs(G1 - G2)) <= 1e-13 * max(np.max(np.abs(G2)), 1e-6)


def test_taylor_method_segmented(lib_built):
    p, eps = configs.random_problem(K=4, N=3, L=2, NT=19, seed=3, gradient_method=gb.TAYLOR)
    check(p, eps)
    # a loose taylor tolerance truncates every chi'_l exactly where taylor_grad_step! returns
    # (per control, optimize.jl:633-638) and must not affect the propagation of chi itself
    p, eps = configs.random_problem(K=4, N=3, L=2, NT=19, seed=3, gradient_method=gb.TAYLOR,
                                    taylor_tolerance=1e-6)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    _, J, G = run(p, eps)
    assert np.max(np.abs(G - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    p2, _ = configs.random_problem(K=4, N=3, L=2, NT=19, seed=3, gradient_method=gb.TAYLOR,
                                   taylor_tolerance=1e-6, path=gb.PATH_SMALL_CHAIN)
    _, J2, G2 = run(p2, eps)
    assert np.max(np.abs(G2 - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    p3, eps3 = configs.random_problem(K=3, N=7, L=2, NT=11, seed=5, gradient_method=gb.TAYLOR,
                                     taylor_tolerance=1e-6)
    ref3 = go.evaluate_gradient(go.from_problem(p3), eps3)
    _, J3, G3 = run(p3, eps3)
    assert np.max(np.abs(G3 - ref3["G"])) <= RTOL * np.max(np.abs(ref3["G"]))


def test_functional_only_then_stored_states(lib_built):
    """evaluate_functional skips the interior fill; stored states are produced on demand."""
    p, eps = configs.c3_ensemble(n_delta=3, n_amp=3, NT=90)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    e = engine(p)
    J = e.evaluate_functional(eps)
    assert abs(J - ref["J"]) <= RTOL
    assert np.max(np.abs(e.stored_states(4) - ref["storage"][4])) <= 1e-12
    G = np.zeros_like(eps)
    e.evaluate_gradient(G, eps)
    assert np.max(np.abs(G - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))


def test_host_chi_segmented(lib_built):
    p, eps = configs.c3_ensemble(n_delta=4, n_amp=4, NT=75, functional=gb.HOST)
    pr, _ = configs.c3_ensemble(n_delta=4, n_amp=4, NT=75, functional=gb.SS)
    ref = go.evaluate_gradient(go.from_problem(pr), eps)
    e = engine(p)
    e.forward(eps)
    psiT = e.final_states()
    tau = np.einsum("ki,ki->k", p.tgt.conj(), psiT)
    Gp = np.zeros_like(eps)
    e.backward_chi((tau / p.K)[:, None] * p.tgt, Gp)
    assert np.max(np.abs(Gp - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))


def test_c3_full_size_properties(lib_built):
    """BASELINE configs[2] at full size: unitarity of the stored states, J_T_ss bounds,
    gradient finite and deterministic, agreement with the plain-chain schedule."""
    p, eps = configs.c3_ensemble()
    e, J, G = run(p, eps)
    assert 0.0 <= J <= 1.0 and np.all(np.isfinite(G))
    for k in (0, 1777, 4095):
        st = e.stored_states(k)
        assert np.max(np.abs(np.sum(np.abs(st) ** 2, axis=0) - 1.0)) < 1e-12
    assert abs(J - (1.0 - np.mean(np.abs(e.tau_vals) ** 2))) < 1e-13
    G2 = np.zeros_like(eps)
    e.evaluate_gradient(G2, eps)
    assert np.array_equal(G, G2)
    e.close()
    p2, _ = configs.c3_ensemble(path=gb.PATH_SMALL_CHAIN)
    _, J2, Gc = run(p2, eps)
    assert abs(J - J2) < 1e-13
    assert np.max(np.abs(G - Gc)) <= 1e-12 * np.max(np.abs(Gc))


# ---------------------------------------------------------------- sub-warp path (5 <= N <= 32)
@pytest.mark.parametrize("seg_len", [2, 3, 7, 128], indirect=True)
@pytest.mark.parametrize("N", [5, 6, 9, 16, 17, 32])
def test_warp_segment_lengths(lib_built, seg_len, N):
    p, eps = configs.random_problem(K=5, N=N, L=2, NT=23, seed=200 + N, hermitian=False, shaped=True,
                                    functional=gb.SM, G=2)
    p.tlist[:] = p.tlist * (0.6 / np.sqrt(N))
    check(p, eps)


def test_warp_matches_plain_chain_schedule(lib_built):
    p, eps = configs.c2_transmon(NT=300)
    e, J1, G1 = run(p, eps)
    p2, _ = configs.c2_transmon(NT=300, path=gb.PATH_WARP_CHAIN)
    _, J2, G2 = run(p2, eps)
    assert abs(J1 - J2) <= 1e-13
    assert np.max(np.abs(G1 - G2)) <= 1e-12 * np.max(np.abs(G2))
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    assert np.max(np.abs(G1 - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    chi, rho = e.chi_states()
    assert np.max(np.abs(chi - ref["chi_states"])) < 1e-12


def test_warp_functional_only_and_host_chi(lib_built):
    p, eps = configs.c2_transmon(NT=120)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    e = engine(p)
    assert abs(e.evaluate_functional(eps) - ref["J"]) <= RTOL
    assert np.max(np.abs(e.stored_states(2) - ref["storage"][2])) <= 1e-12
    ph, _ = configs.c2_transmon(NT=120, functional=gb.HOST)
    eh = engine(ph)
    eh.forward(eps)
    tau = np.einsum("ki,ki->k", ph.tgt.conj(), eh.final_states())
    chi = (np.sum(tau) / ph.K ** 2) * ph.tgt          # J_T_sm: chi_k = (sum_j tau_j / K^2) |tgt_k>
    Gp = np.zeros_like(eps)
    eh.backward_chi(chi, Gp)
    assert np.max(np.abs(Gp - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))


# ---- Hermitian schedule: forward states recomputed backwards inside the gradient kernel (csrc/small_seg.cuh,
# ---- seg_step_krylov_h), fw_storage filled lazily; GRAPE_B200_SEG_HERM=0 selects the general schedule
@pytest.fixture
def herm_mode(request):
    old = os.environ.get("GRAPE_B200_SEG_HERM")
    os.environ["GRAPE_B200_SEG_HERM"] = str(request.param)
    yield request.param
    if old is None:
        del os.environ["GRAPE_B200_SEG_HERM"]
    else:
        os.environ["GRAPE_B200_SEG_HERM"] = old


@pytest.mark.parametrize("herm_mode", [1, 0], indirect=True)
@pytest.mark.parametrize("seg_len", [2, 7, 64], indirect=True)
@pytest.mark.parametrize("N", [1, 2, 3])
def test_hermitian_schedule_small_theta(lib_built, herm_mode, seg_len, N):
    p, eps = configs.random_problem(K=5, N=N, L=2, NT=37, seed=300 + N, hermitian=True, shaped=True, functional=gb.SM)
    p.tlist[:] = p.tlist * 0.02
    check(p, eps)


@pytest.mark.parametrize("herm_mode", [1, 0], indirect=True)
@pytest.mark.parametrize("K", [1, 3, 33, 70])
def test_hermitian_schedule_lane_mapping_and_large_theta(lib_built, herm_mode, K):
    """large ||H dt|| (sub-steps, block recursion in the gradient kernel) with Hermitian generators"""
    p, eps = configs.random_problem(K=K, N=3, L=3, NT=19, seed=310 + K, hermitian=True, uniform=True, functional=gb.SS)
    check(p, eps)
    p, eps = configs.random_problem(K=K, N=2, L=1, NT=23, seed=320 + K, hermitian=True, functional=gb.RE)
    p.tlist[:] = p.tlist * 0.05
    check(p, eps)


@pytest.mark.parametrize("herm_mode", [1, 0], indirect=True)
def test_hermitian_schedule_taylor_method(lib_built, herm_mode):
    p, eps = configs.random_problem(K=4, N=3, L=2, NT=21, seed=331, hermitian=True, gradient_method=gb.TAYLOR)
    p.tlist[:] = p.tlist * 0.1
    check(p, eps)


def test_hermitian_schedule_lazy_storage_follows_the_pulses(lib_built):
    """fw_storage is filled on demand from the pulses of the LAST evaluation (also after a graph replay and after
    the device-pointer entry points)"""
    import torch
    p, eps = configs.c3_ensemble(n_delta=6, n_amp=5, NT=90)
    op = go.from_problem(p)
    e = engine(p)
    G = np.zeros_like(eps)
    for scale in (1.0, 0.6, 1.3):
        x = eps * scale
        ref = go.evaluate_gradient(op, x)
        J = e.evaluate_gradient(G, x)
        assert abs(J - ref["J"]) <= RTOL and np.max(np.abs(G - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
        assert np.max(np.abs(e.stored_states(7) - ref["storage"][7])) <= 1e-12
        assert np.max(np.abs(e.final_states() - ref["final_states"])) <= 1e-12
    x = eps * 0.8
    ref = go.evaluate_gradient(op, x)
    d_x = torch.from_numpy(x).cuda()
    d_G = torch.zeros_like(d_x)
    e.eval_fg_device(d_x.data_ptr(), d_G.data_ptr(), None)
    assert np.max(np.abs(d_G.cpu().numpy() - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    assert np.max(np.abs(e.stored_states(11) - ref["storage"][11])) <= 1e-12
    assert np.max(np.abs(e.tau_grads(11) - ref["tau_grads"][11])) <= 1e-12
    e.close()


def test_c3_full_size_hermitian_vs_general_schedule(lib_built):
    """BASELINE configs[2] at full size: the two schedules agree to 1e-12 on every gradient element"""
    p, eps = configs.c3_ensemble()
    e, J, G = run(p, eps)
    e.close()
    os.environ["GRAPE_B200_SEG_HERM"] = "0"
    try:
        e0, J0, G0 = run(p, eps)
        e0.close()
    finally:
        del os.environ["GRAPE_B200_SEG_HERM"]
    assert abs(J - J0) < 1e-13
    assert np.max(np.abs(G - G0)) <= 1e-12 * np.max(np.abs(G0))


# ---- sub-warp path, Hermitian generators: scan schedule (csrc/warp_seg.cuh warp_segscan / warp_scan_bounds_*);
# ---- GRAPE_B200_WSEG_SCAN=0 keeps the boundary chains
@pytest.fixture
def wscan(request):
    old = os.environ.get("GRAPE_B200_WSEG_SCAN")
    os.environ["GRAPE_B200_WSEG_SCAN"] = str(request.param)
    yield request.param
    if old is None:
        del os.environ["GRAPE_B200_WSEG_SCAN"]
    else:
        os.environ["GRAPE_B200_WSEG_SCAN"] = old


@pytest.mark.parametrize("wscan", [1, 0], indirect=True)
@pytest.mark.parametrize("N,NT", [(5, 16), (6, 37), (9, 300), (16, 64), (17, 100), (32, 36)])   # the Python oracle sets the cost
def test_warp_scan_schedule_hermitian(lib_built, wscan, N, NT):
    """prefix products over the segments by a scan, boundary states Psi = Q_seg Psi(0), chi = Q_seg Q_last^dagger chi(T)
    (every P unitary): every sub-warp width, ragged last segment, two generators, shaped pulses, all three functionals"""
    for fn in (gb.SM, gb.RE, gb.SS):
        p, eps = configs.random_problem(K=5, N=N, L=2, NT=NT, seed=600 + N + fn, hermitian=True, shaped=True,
                                        functional=fn, G=2, weights=np.linspace(0.5, 1.5, 5))
        p.tlist[:] = p.tlist * (0.6 / np.sqrt(N))
        check(p, eps)[0].close()


@pytest.mark.parametrize("seg_len", [2, 3, 5], indirect=True)
@pytest.mark.parametrize("N", [6, 20, 32])
def test_warp_scan_more_segments_than_sub_warps(lib_built, seg_len, N):
    """forced short segments: many segments per generator, several scan levels of both radices"""
    NT = 300 if N == 6 else (120 if N == 20 else 60)
    p, eps = configs.random_problem(K=2, N=N, L=2, NT=NT, seed=640 + N, hermitian=True, functional=gb.SM, G=1)
    p.tlist[:] = p.tlist * (0.6 / np.sqrt(N))
    check(p, eps)[0].close()


def test_warp_scan_matches_chain_schedule_and_host_chi(lib_built):
    p, eps = configs.c2_transmon(NT=500)
    out = {}
    for mode in (1, 0):
        os.environ["GRAPE_B200_WSEG_SCAN"] = str(mode)
        try:
            e, J, G = run(p, eps)
        finally:
            del os.environ["GRAPE_B200_WSEG_SCAN"]
        chi, rho = e.chi_states()
        out[mode] = (J, G, chi, rho, e.stored_states(1))
        e.close()
    assert abs(out[1][0] - out[0][0]) <= 1e-13
    assert np.max(np.abs(out[1][1] - out[0][1])) <= 1e-12 * np.max(np.abs(out[0][1]))
    assert np.max(np.abs(out[1][2] - out[0][2])) <= 1e-12 and np.max(np.abs(out[1][3] - out[0][3])) <= 1e-12
    assert np.max(np.abs(out[1][4] - out[0][4])) <= 1e-12
    # host-supplied chi through the scan schedule
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    ph, _ = configs.c2_transmon(NT=500, functional=gb.HOST)
    eh = engine(ph)
    eh.forward(eps)
    tau = np.einsum("ki,ki->k", ph.tgt.conj(), eh.final_states())
    Gp = np.zeros_like(eps)
    eh.backward_chi((np.sum(tau) / ph.K ** 2) * ph.tgt, Gp)
    assert np.max(np.abs(Gp - ref["G"])) <= RTOL * np.max(np.abs(ref["G"]))
    eh.close()
