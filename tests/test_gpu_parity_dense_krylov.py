"""GPU parity of the two backward forms of the dense path against the CPU oracle and against each other.

Krylov form (csrc/dense_kry.cuh; default when no step needs sub-stepping): chi chain on K columns, Taylor
terms of both sweeps kept in HBM, one DMMA contraction per time step for all controls.
Block recursion (csrc/dense.cuh, dense2.cuh; GRAPE_B200_KRYLOV=0, sub-stepped steps, :taylor): the
GradGenerator block vector of the reference (src/optimize.jl:880-911) propagated term by term.
Tolerance 1e-10 relative on J and every gradient element (north_star)."""
import os

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import grape_oracle as go
from tests.test_gpu_parity_small import check, engine

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _both_forms(p, eps, rtol=1e-10, **env):
    """oracle parity (J, J_parts, tau, G, grad_J_Tb, grad_J_a, evaluate_functional) with the Krylov form on
    (default) and off; the two gradients agree to 1e-11. Returns them."""
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    scale = max(np.max(np.abs(ref["G"])), 1e-6)
    out = []
    for kry, form in ((1, 1), (0, 0)):
        with _Env(GRAPE_B200_KRYLOV=kry, **env):
            e = engine(p)
        G = np.zeros_like(eps)
        J = e.evaluate_gradient(G, eps)
        assert (e.gradient_form() > 0) == bool(form), (kry, e.gradient_form())
        assert abs(J - ref["J"]) <= rtol * max(1.0, abs(ref["J"])), (J, ref["J"])
        assert np.max(np.abs(e.J_parts - ref["J_parts"])) <= rtol * max(1.0, np.max(np.abs(ref["J_parts"])))
        assert np.max(np.abs(e.tau_vals - ref["tau"])) <= rtol
        err = np.max(np.abs(G - ref["G"])) / scale
        assert err <= rtol, f"gradient rel err {err:.3e} (krylov={kry})"
        assert np.max(np.abs(e.grad_J_Tb - ref["grad_J_Tb"])) / scale <= rtol
        assert np.max(np.abs(e.grad_J_a - ref["grad_J_a"])) <= rtol * max(1.0, np.max(np.abs(ref["grad_J_a"])))
        assert abs(e.evaluate_functional(eps) - ref["J"]) <= rtol * max(1.0, abs(ref["J"]))
        out.append(G)
        e.close()
    assert np.max(np.abs(out[0] - out[1])) <= 1e-11 * scale
    return out


@pytest.mark.parametrize("N,K", [(33, 3), (40, 8), (64, 5), (100, 16), (130, 9), (65, 17)])
def test_strip_kernels_both_forms(lib_built, N, K):
    p, eps = configs.c4_dense450(N=N, K=K, NT=6)
    _both_forms(p, eps, GRAPE_B200_DENSE2=0)


@pytest.mark.parametrize("N,K", [(33, 9), (64, 16), (100, 12), (130, 27), (40, 32)])
def test_tiled_kernels_both_forms(lib_built, N, K):
    p, eps = configs.c4_dense450(N=N, K=K, NT=5)
    _both_forms(p, eps, GRAPE_B200_DENSE2=1)


@pytest.mark.parametrize("dense2", [0, 1])
@pytest.mark.parametrize("functional", [gb.SM, gb.RE, gb.SS])
@pytest.mark.parametrize("L", [1, 3, 4])
def test_nonhermitian_shaped_weighted(lib_built, dense2, functional, L):
    N, K = 36, 11
    p, eps = configs.random_problem(K=K, N=N, L=L, NT=5, G=1, seed=270 + functional + 10 * L, hermitian=False,
                                    shaped=True, weights=np.linspace(0.5, 1.5, K), functional=functional)
    p.tlist = p.tlist * (0.5 / np.sqrt(N))
    _both_forms(p, eps, GRAPE_B200_DENSE2=dense2)


@pytest.mark.parametrize("dense2", [0, 1])
def test_running_costs_both_forms(lib_built, dense2):
    """state running cost g_b = <Psi|D|Psi> (chi inhomogeneity in the chain, optimize.jl:897-908) + fluence J_a"""
    p, eps = configs.c5_dense1024(N=40, K=16, NT=7)
    _both_forms(p, eps, GRAPE_B200_DENSE2=dense2)
    # non-uniform time grid: trapezoid weights and Taylor orders differ per step
    p, eps = configs.c5_dense1024(N=48, K=8, NT=6)
    p.tlist = np.cumsum(np.concatenate([[0.0], 0.25 * (1.0 + 0.4 * np.sin(1.0 + np.arange(p.NT)))]))
    _both_forms(p, eps, GRAPE_B200_DENSE2=dense2)


@pytest.mark.parametrize("dense2", [0, 1])
def test_substeps_fall_back_to_block_recursion(lib_built, dense2):
    """||H dt|| ~ 4 needs sub-steps: the device-side plan selects the block recursion for the call"""
    p, eps = configs.c4_dense450(N=48, K=8, NT=4)
    p.tlist = p.tlist * 8.0
    with _Env(GRAPE_B200_DENSE2=dense2):
        e, ref = check(p, eps, rtol=1e-9)
    assert e.gradient_form() == 0
    e.close()


def test_form_switches_per_call_on_one_handle(lib_built):
    """one handle, pulse vectors with ||H dt|| <= 1 on every step (Krylov form) and > 1 on some steps (block
    recursion), alternating: the device-side plan picks the kernels per call and both match the oracle"""
    p, eps = configs.c4_dense450(N=48, K=8, NT=5)
    p.tlist = p.tlist * 2.5
    e = engine(p)
    op = go.from_problem(p)
    for scale, form in ((0.1, 1), (1.0, 0), (0.1, 1), (1.0, 0)):
        x = eps * scale
        ref = go.evaluate_gradient(op, x)
        G = np.zeros_like(x)
        J = e.evaluate_gradient(G, x)
        assert (e.gradient_form() > 0) == bool(form)
        assert abs(J - ref["J"]) <= 1e-9
        assert np.max(np.abs(G - ref["G"])) <= 1e-9 * max(np.max(np.abs(ref["G"])), 1e-6)
    e.close()


def test_few_term_slots_fall_back(lib_built):
    """GRAPE_B200_KRY_MT smaller than the Taylor order of the steps -> block recursion, same numbers"""
    p, eps = configs.c4_dense450(N=40, K=8, NT=4)
    with _Env(GRAPE_B200_KRY_MT=6):
        e, ref = check(p, eps)
        assert e.gradient_form() == 0
    e.close()


def test_taylor_method_uses_block_recursion(lib_built):
    p, eps = configs.c4_dense450(N=40, K=8, NT=4, gradient_method=gb.TAYLOR)
    e, ref = check(p, eps)
    assert e.gradient_form() == 0
    e.close()


def test_split_forward_backward_and_host_chi(lib_built):
    """sharded-style split call and the host-functional round trip go through the Krylov form too"""
    p, eps = configs.c4_dense450(N=50, K=7, NT=5)
    e, ref = check(p, eps)
    G = np.zeros_like(eps)
    e.evaluate_gradient(G, eps)
    sums = e.forward(eps)
    Gp = np.zeros_like(eps)
    e.backward(sums, Gp)
    assert e.gradient_form() >= 1
    # eval_fg runs the two chains concurrently (chi_k(T) applied in the contraction), the split call one after the
    # other: same series, different rounding
    assert np.max(np.abs(Gp - G)) <= 1e-12 * np.max(np.abs(G))
    e.close()
    ph, _ = configs.c4_dense450(N=50, K=7, NT=5)
    ph.functional = gb.HOST
    eh = engine(ph)
    eh.forward(eps)
    psiT = eh.final_states()
    tau = np.sum(np.conj(ph.tgt) * psiT, axis=1)
    chiT = (np.sum(tau) / ph.K ** 2) * ph.tgt          # chi of J_T_sm, evaluated on the host
    Gh = np.zeros_like(eps)
    eh.backward_chi(chiT, Gh)
    assert eh.gradient_form() >= 1
    assert np.max(np.abs(Gh - ref["G"])) <= 1e-10 * np.max(np.abs(ref["G"]))
    eh.close()


def test_c4_full_width_forms_agree(lib_built):
    """C4 at full width (N=450, K=16), 3 steps: oracle parity of the Krylov form, block recursion within 1e-11"""
    p, eps = configs.c4_dense450(NT=3)
    _both_forms(p, eps)


def test_c5_full_width_forms_agree(lib_built):
    """C5 at full width (N=1024, K=64, J_a + g_b), 2 steps: Krylov form vs block recursion (the dense Pade oracle
    of a 3072 x 3072 block matrix per trajectory-step is too slow for 64 trajectories)"""
    p, eps = configs.c5_dense1024(NT=2)
    res = []
    for kry in (1, 0):
        with _Env(GRAPE_B200_KRYLOV=kry):
            e = engine(p)
        G = np.zeros_like(eps)
        J = e.evaluate_gradient(G, eps)
        assert (e.gradient_form() > 0) == bool(kry)
        res.append((J, G))
        e.close()
    assert abs(res[0][0] - res[1][0]) <= 1e-12
    assert np.max(np.abs(res[0][1] - res[1][1])) <= 1e-11 * np.max(np.abs(res[1][1]))


def test_large_n_few_trajectories_uses_tiled_kernels(lib_built):
    """N = 1216 > 1184: more 8-row strips than SMs, so the strip kernels cannot run; the trajectory block is padded
    to the 16-column tile and the tiled kernels serve both backward forms.  Oracle: the :taylor variant (N x N
    exponentials only), which agrees with :gradgen to 1e-14 (tests/test_oracle_known_answers.py)."""
    p, eps = configs.c4_dense450(N=1216, K=3, NT=2)
    ref = go.evaluate_gradient(go.from_problem(p, gradient_method=go.TAYLOR), eps)
    scale = np.max(np.abs(ref["G"]))
    for kry in (1, 0):
        with _Env(GRAPE_B200_KRYLOV=kry):
            e = engine(p)
        G = np.zeros_like(eps)
        J = e.evaluate_gradient(G, eps)
        assert (e.gradient_form() > 0) == bool(kry)
        assert abs(J - ref["J"]) <= 1e-10
        assert np.max(np.abs(G - ref["G"])) <= 1e-10 * scale
        assert np.max(np.abs(e.final_states() - ref["final_states"])) <= 1e-12
        e.close()


@pytest.mark.parametrize("dense2", [0, 1])
@pytest.mark.parametrize("N,K,NT,L", [(450, 24, 3, 2), (450, 16, 4, 2), (100, 16, 6, 3), (130, 9, 5, 1), (64, 40, 5, 2)])
def test_several_terms_per_barrier_match_single_term_chain(lib_built, N, K, NT, L, dense2):
    """chains with operand tiles of H_n^2 (and H_n^3): two / three Taylor terms per grid barrier (strip kernels
    csrc/dense.cuh dense_chain<BWD, NS> -- one or several 8-column groups per CTA -- and tiled kernels csrc/dense2.cuh
    dense2_chain_multi<BWD, NS>; Taylor orders of every residue mod NS, 1..3 controls, Hermitian and non-Hermitian
    generators) against the one-term-per-barrier chain: same truncated series, so J and every gradient element agree
    to 1e-12"""
    if dense2 and ((K + 7) // 8 * 8) % 16 != 0:
        pytest.skip("trajectory block is not a multiple of the 16-column tile: the strip kernels serve this shape")
    if L == 2:
        p, eps = configs.c4_dense450(N=N, K=K, NT=NT)
    else:
        p, eps = configs.random_problem(K=K, N=N, L=L, NT=NT, G=1, seed=500 + N, hermitian=False, shaped=True)
        p.tlist = p.tlist * (0.5 / np.sqrt(N))
    res = []
    # three / two terms per barrier with the generators of all steps pre-formed once per call (tiles of H_n, H_n^2
    # [, H_n^3] fetched by cp.async; the non-Hermitian cases keep a second set for the adjoints), two terms with the
    # strip chain forming its strips itself (the tiled chain has no such mode: one term), and the single-term chain
    modes = ((3, 1, 2), (2, 1, 2), (2, 0, 1 if dense2 else 2), (1, 0, 1))
    for terms, pre, form in modes:
        with _Env(GRAPE_B200_DENSE2=dense2, GRAPE_B200_DENSE_TERMS=terms, GRAPE_B200_DENSE_PREFORM=pre):
            e = engine(p)
        G = np.zeros_like(eps)
        for _ in range(2):                       # twice: the prefetch state must not leak between calls
            J = e.evaluate_gradient(G, eps)
        assert e.gradient_form() == form, (terms, pre, e.gradient_form())
        res.append((J, G.copy(), e.final_states(), e.stored_states(K - 1)))
        e.close()
    scale = np.max(np.abs(res[3][1]))
    for r in res[:3]:
        assert abs(r[0] - res[3][0]) <= 1e-12
        assert np.max(np.abs(r[1] - res[3][1])) <= 1e-12 * scale
        assert np.max(np.abs(r[2] - res[3][2])) <= 1e-13
        assert np.max(np.abs(r[3] - res[3][3])) <= 1e-13
    if not dense2:
        assert np.array_equal(res[1][1], res[2][1])   # same arithmetic, only the operand source differs


def test_two_terms_per_barrier_full_width_oracle(lib_built):
    """N = 450 with three 8-column groups on two column parts (1 and 2 groups per CTA): oracle parity of the dual chain
    (:taylor oracle = N x N exponentials only; agrees with :gradgen to 1e-14, tests/test_oracle_known_answers.py)"""
    p, eps = configs.c4_dense450(N=450, K=24, NT=2)
    ref = go.evaluate_gradient(go.from_problem(p, gradient_method=go.TAYLOR), eps)
    e = engine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    assert e.gradient_form() == 2
    assert abs(J - ref["J"]) <= 1e-10
    assert np.max(np.abs(G - ref["G"])) <= 1e-10 * np.max(np.abs(ref["G"]))
    assert np.max(np.abs(e.final_states() - ref["final_states"])) <= 1e-12
    e.close()


@pytest.mark.parametrize("functional,terms", [(gb.SM, 3), (gb.SS, 3), (gb.RE, 2), (gb.SM, 1)])
def test_concurrent_forward_and_backward_chains(lib_built, functional, terms):
    """dense_chain<2, NS> (csrc/dense.cuh): the forward sweep and the chi chain run at the same time in one cooperative
    grid, chi_k(T) = c_k tgt_k being applied afterwards in the contraction.  Against the oracle, against the sequential
    sweeps (GRAPE_B200_DENSE_CONCURRENT=0), with Taylor orders that differ between step n and step NT-1-n (pulse ramp:
    the two directions need different numbers of grid barriers per iteration), weights, and call after call."""
    N, K, NT = 48, 16, 9
    w = np.linspace(0.5, 1.5, K)
    p, eps = configs.c4_dense450(N=N, K=K, NT=NT, functional=functional, weights=w)
    ramp = np.linspace(0.05, 2.5, NT)                      # ||H_n dt|| from ~0.1 to ~0.9: orders 9 .. 18
    eps = (eps.reshape(2, NT) * ramp[None, :]).reshape(-1)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    scale = np.max(np.abs(ref["G"]))
    res = {}
    for conc in (1, 0):
        with _Env(GRAPE_B200_DENSE2=0, GRAPE_B200_DENSE_CONCURRENT=conc, GRAPE_B200_DENSE_TERMS=terms):
            e = engine(p)
        G = np.zeros_like(eps)
        for rep in range(2):
            J = e.evaluate_gradient(G, eps)
            assert e.dense_concurrent() == conc and e.gradient_form() > 0
            assert abs(J - ref["J"]) <= 1e-10 and np.max(np.abs(e.tau_vals - ref["tau"])) <= 1e-10
            assert np.max(np.abs(G - ref["G"])) <= 1e-10 * scale
            chi, rho = e.chi_states()
            assert np.max(np.abs(chi - ref["chi_states"])) <= 1e-10 and np.max(np.abs(rho - ref["chi_norms"])) <= 1e-12
            assert abs(e.evaluate_functional(eps) - ref["J"]) <= 1e-10          # forward-only call in between
        # split host API: forward, then backward with the (here: local = global) sums -> sequential sweeps
        sums = e.forward(eps)
        Gp = np.zeros_like(eps)
        e.backward(sums, Gp)
        assert e.dense_concurrent() == 0
        assert np.max(np.abs(Gp - ref["grad_J_Tb"])) <= 1e-10 * scale
        # sub-stepped pulses (||H dt|| > 1): the block recursion serves the call, then the concurrent chains again
        big = eps * 6.0
        refb = go.evaluate_gradient(go.from_problem(p), big)
        Jb = e.evaluate_gradient(G, big)
        assert e.gradient_form() == 0 and e.dense_concurrent() == 0
        assert abs(Jb - refb["J"]) <= 1e-9 and np.max(np.abs(G - refb["G"])) <= 1e-9 * np.max(np.abs(refb["G"]))
        J = e.evaluate_gradient(G, eps)
        assert e.dense_concurrent() == conc and np.max(np.abs(G - ref["G"])) <= 1e-10 * scale
        res[conc] = G.copy()
        e.close()
    assert np.max(np.abs(res[0] - res[1])) <= 1e-12 * scale


def test_concurrent_chains_not_used_with_state_running_cost_or_host_chi(lib_built):
    p, eps = configs.c5_dense1024(N=40, K=8, NT=6)                         # g_b: chi needs Psi(t_{n-1}) on its way back
    with _Env(GRAPE_B200_DENSE2=0):
        e = engine(p)
    G = np.zeros_like(eps)
    e.evaluate_gradient(G, eps)
    assert e.dense_concurrent() == 0 and e.gradient_form() > 0
    e.close()
    p, eps = configs.c4_dense450(N=40, K=8, NT=6, functional=gb.HOST)
    pr, _ = configs.c4_dense450(N=40, K=8, NT=6, functional=gb.SS)
    ref = go.evaluate_gradient(go.from_problem(pr), eps)
    with _Env(GRAPE_B200_DENSE2=0):
        e = engine(p)
    e.forward(eps)
    tau = np.einsum("ki,ki->k", p.tgt.conj(), e.final_states())
    Gp = np.zeros_like(eps)
    e.backward_chi((tau / p.K)[:, None] * p.tgt, Gp)
    assert e.dense_concurrent() == 0
    assert np.max(np.abs(Gp - ref["G"])) <= 1e-10 * np.max(np.abs(ref["G"]))
    e.close()


@pytest.mark.parametrize("conc", [1, 0])
def test_tma_strip_prefetch_matches_cp_async(lib_built, conc):
    """operator strips of the multi-term chains through the TMA copy engine (cp.async.bulk + mbarrier, default) or the
    per-thread cp.async path (GRAPE_B200_DENSE_TMA=0): bit-identical gradients, both at 1e-10 of the oracle"""
    p, eps = configs.c4_dense450(N=96, K=16, NT=11)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    out = []
    for tma in (1, 0):
        with _Env(GRAPE_B200_DENSE2=0, GRAPE_B200_DENSE_TMA=tma, GRAPE_B200_DENSE_CONCURRENT=conc):
            e = engine(p)
        G = np.zeros_like(eps)
        for _ in range(2):
            J = e.evaluate_gradient(G, eps)
        assert e.gradient_form() == 2 and e.dense_concurrent() == conc
        assert abs(J - ref["J"]) <= 1e-10 and np.max(np.abs(G - ref["G"])) <= 1e-10 * np.max(np.abs(ref["G"]))
        out.append(G.copy())
        e.close()
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("dense2", [0, 1])
def test_economised_polynomial_orders_and_parity(lib_built, dense2):
    """Hermitian generators: the Krylov-form chains sum the Chebyshev-cut polynomial of exp(-i H dt) (csrc/dense.cuh
    econ_table; the reference's Cheby propagator, docs/src/tutorial.md:308, 432, in the monomial basis) -- degree <= 12
    where the Taylor series needs 13..16 terms at ||H dt|| <= 0.525 --, the gradient is the exact derivative of that
    polynomial (beta(a,b) g[a+b+1]); both series match the oracle at 1e-10 and each other far below that."""
    p, eps = configs.c5_dense1024(N=64, K=16, NT=12)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    scale = max(np.max(np.abs(ref["G"])), 1e-6)
    got = {}
    for econ in (1, 0):
        with _Env(GRAPE_B200_ECON=econ, GRAPE_B200_DENSE2=dense2):
            e = engine(p)
        G = np.zeros_like(eps)
        J = e.evaluate_gradient(G, eps)
        assert e.gradient_form() > 0
        flag, orders = e.dense_orders()
        assert flag == bool(econ)
        assert abs(J - ref["J"]) <= 1e-10 * max(1.0, abs(ref["J"]))
        assert np.max(np.abs(G - ref["G"])) <= 1e-10 * scale
        assert np.max(np.abs(e.tau_vals - ref["tau"])) <= 1e-10
        assert abs(e.evaluate_functional(eps) - J) <= 1e-13 * max(1.0, abs(J))
        got[econ] = (J, G, orders)
        e.close()
    assert got[1][2].max() <= 12 and got[0][2].min() >= 13, (got[1][2], got[0][2])
    assert abs(got[1][0] - got[0][0]) <= 1e-13 * max(1.0, abs(got[0][0]))
    assert np.max(np.abs(got[1][1] - got[0][1])) <= 1e-12 * scale


def test_economised_polynomial_only_for_hermitian_generators(lib_built):
    p, eps = configs.random_problem(K=8, N=40, L=2, NT=5, G=1, seed=77, hermitian=False)
    p.tlist = p.tlist * (0.5 / np.sqrt(p.N))
    e, ref = check(p, eps)
    assert e.gradient_form() > 0
    flag, orders = e.dense_orders()
    assert flag is False
    e.close()


def test_economised_polynomial_long_chain_unitarity(lib_built):
    """1000 steps of the economised polynomial: the final states keep their norm to 1e-12 (a per-step error of 1e-17
    accumulates linearly at worst) and J matches the Taylor-series run to 1e-12."""
    p, eps = configs.c4_dense450(N=96, K=8, NT=1000)
    vals = {}
    for econ in (1, 0):
        with _Env(GRAPE_B200_ECON=econ):
            e = engine(p)
        vals[econ] = e.evaluate_functional(eps)
        psi = e.final_states()
        assert np.max(np.abs(np.linalg.norm(psi, axis=1) - 1.0)) <= 1e-12
        e.close()
    assert abs(vals[1] - vals[0]) <= 1e-12
