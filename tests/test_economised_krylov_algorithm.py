"""CPU restatement of what the dense Krylov-form kernels compute with the economised series (csrc/econ.cuh,
csrc/dense.cuh dense_chain, csrc/dense_kry.cuh kry_combine / kry_contract), in NumPy, against the exact oracle:

    forward   Psi_n   = sum_{j<=m} g_j bh_j,   bh_j = (-i H dt)^j Psi_{n-1} / j!          (Taylor terms, weights g)
    backward  chi_n-1 = sum_{j<=m} g_j ch_j,   ch_j = (+i H^dagger dt)^j chi_n / j!
    gradient  <chi'_l | Psi> = Tr(E_l^dagger M),  M = sum_{a+b<=m-1} beta(a,b) g_{a+b+1} bh_a ch_b^dagger,
              beta(a,b) = a! b! / (a+b+1)!,  E_l = +i dt mu_l^dagger

with the degree m and the weights g taken from the library's own table (host-only entry point).  This pins the
ALGORITHM -- in particular the claim that beta(a,b) g_{a+b+1} is the exact derivative of the polynomial propagator --
independently of any GPU: J and every gradient element agree with the oracle's dense expm / GradGenerator step
(reference src/optimize.jl:880-911, docs/src/background.md:447-494) to 1e-12."""
import ctypes as C
import math

import numpy as np

from grape.jl_b200 import _lib, configs
from oracle import grape_oracle as go


def _table(lib, m):
    th = C.c_double()
    g = (C.c_double * (m + 1))()
    assert lib.grape_b200_econ_table(m, C.byref(th), g) == 0
    return th.value, np.array(g[:])


def _degree(lib, theta):
    for m in range(2, 21):
        th, g = _table(lib, m)
        if th >= theta:
            return m, g
    raise AssertionError("theta > 1: the kernels would sub-step")


def _economised_gradient(lib, p, eps):
    """J_T_sm functional, no running costs: the schedule of the kernels, one trajectory block at a time"""
    K, N, L, NT = p.K, p.N, p.L, p.NT
    e = eps.reshape(L, NT)
    H0, Hc = p.H0[0], p.Hc[0]
    hn = [1.05 * np.linalg.norm(H0, 2)] + [1.05 * np.linalg.norm(Hc[l], 2) for l in range(L)]   # dense_setup
    dts = np.diff(p.tlist)
    psi = p.psi0.T.copy()                                # [N, K]
    fw_terms, orders = [], []
    for n in range(NT):
        H = H0 + sum(e[l, n] * Hc[l] for l in range(L))
        theta = (hn[0] + sum(abs(e[l, n]) * hn[1 + l] for l in range(L))) * dts[n]
        m, g = _degree(lib, theta)
        terms = [psi]
        for j in range(1, m + 1):
            terms.append((-1j * dts[n] / j) * (H @ terms[-1]))
        psi = sum(g[j] * terms[j] for j in range(m + 1))
        fw_terms.append(terms)
        orders.append((m, g))
    tau = np.einsum("ki,ik->k", p.tgt.conj(), psi)
    J = 1.0 - abs(np.sum(tau)) ** 2 / K ** 2
    chi = (np.sum(tau) / K ** 2) * p.tgt.T               # chi_k = (sum_j tau_j / K^2) tgt_k   (optimize.jl:845-855)
    rho = np.linalg.norm(chi, axis=0)
    chi = chi / rho
    G = np.zeros((L, NT))
    for n in range(NT - 1, -1, -1):
        m, g = orders[n]
        H = H0 + sum(e[l, n] * Hc[l] for l in range(L))
        ch = [chi]
        for j in range(1, m + 1):
            ch.append((1j * dts[n] / j) * (H.conj().T @ ch[-1]))
        bh = fw_terms[n]
        for k in range(K):
            M = np.zeros((N, N), dtype=np.complex128)
            for a in range(m):
                for b in range(m - a):
                    beta = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 1)
                    M += beta * g[a + b + 1] * np.outer(bh[a][:, k], ch[b][:, k].conj())
            for l in range(L):
                E = 1j * dts[n] * Hc[l].conj().T
                G[l, n] += -2.0 * rho[k] * np.real(np.sum(E.conj() * M))      # optimize.jl:574-584, 893-895
        chi = sum(g[j] * ch[j] for j in range(m + 1))
    return J, G.reshape(-1), orders


def test_economised_krylov_form_matches_the_exact_oracle(lib_built):
    lib = _lib.load()
    p, eps = configs.c4_dense450(N=12, K=3, NT=7)
    J, G, orders = _economised_gradient(lib, p, eps)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    assert abs(J - ref["J"]) <= 1e-13
    assert np.max(np.abs(G - ref["G"])) <= 1e-12 * max(np.max(np.abs(ref["G"])), 1e-6)
    # the same steps with the Taylor weights (g = 1) at the economised degree miss the oracle by far more: the
    # weights, not the degree, carry the accuracy
    assert max(m for m, _ in orders) <= 12


def test_taylor_weights_at_the_economised_degree_are_not_enough(lib_built):
    lib = _lib.load()
    theta = 0.5
    m, g = _degree(lib, theta)
    x = np.linspace(-theta, theta, 201)
    econ = sum(g[j] * (-1j * x) ** j / math.factorial(j) for j in range(m + 1))
    tayl = sum((-1j * x) ** j / math.factorial(j) for j in range(m + 1))
    assert np.max(np.abs(econ - np.exp(-1j * x))) <= 5e-16          # double rounding of this check itself
    assert np.max(np.abs(tayl - np.exp(-1j * x))) >= 1e-14
