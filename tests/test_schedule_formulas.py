"""CPU checks (NumPy, no GPU) of the algebra behind two device schedules, against the oracle:

* csrc/small_sym.cuh -- real-symmetric generators: U_n = cos(Hs) - i sin(Hs) from the even / odd halves of the
  degree-d Taylor polynomial in Q = Hs^2, and the end-of-step Krylov form of the gradient with un-normalised
  real-matrix powers, kappa(a,b) = 1/((a+b+1) a! b!) and Im M only;
* csrc/dense.cuh / dense2.cuh -- several Taylor terms per grid barrier: H_n^2 and H_n^3 as quadratic / cubic forms
  of the symmetrised operator products, T_{j+q} = (-i dt)^q/((j+1)..(j+q)) H^q T_j.

These are line-by-line NumPy transcriptions of what the kernels compute, so a formula error shows up here without
a GPU; the kernels themselves are compared with the oracle in tests/test_gpu_parity_*.py."""
import math

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs
from oracle import grape_oracle as go

FACT = [float(math.factorial(j)) for j in range(24)]
TH = [0, 0, 6.324548995781438e-09, 4.932419216236795e-06, 0.00014801641288189615, 0.001191356706809193,
      0.004932419216236793, 0.013910766804023767, 0.030783527232167155]      # c_sym_th of small_sym.cuh


def kappa(a, b):
    return 1.0 / ((a + b + 1) * FACT[a] * FACT[b])


def rot(w, r):
    return (1j ** (r & 3)) * w


def cos_sin_poly(Hs, degree):
    """small_formseg_sym: Horner in Q = Hs^2 for both halves of the degree-`degree` Taylor polynomial"""
    N = Hs.shape[0]
    Q = Hs @ Hs
    d2 = (degree - 1) // 2
    sg = -1.0 if d2 & 1 else 1.0
    C = sg / FACT[2 * d2] * Q - sg / FACT[2 * d2 - 2] * np.eye(N)
    Sp = sg / FACT[2 * d2 + 1] * Q - sg / FACT[2 * d2 - 1] * np.eye(N)
    for j in range(d2 - 2, -1, -1):
        sj = -1.0 if j & 1 else 1.0
        C = Q @ C + sj / FACT[2 * j] * np.eye(N)
        Sp = Q @ Sp + sj / FACT[2 * j + 1] * np.eye(N)
    return C, Hs @ Sp


@pytest.mark.parametrize("theta,degree", [(1e-4, 3), (0.03, 7), (0.2, 11), (0.6, 15)])
def test_cos_sin_halves_of_the_taylor_polynomial(theta, degree):
    from scipy.linalg import expm
    rng = np.random.default_rng(3)
    A = rng.standard_normal((3, 3))
    Hs = (A + A.T) / 2
    Hs *= theta / np.max(np.sum(np.abs(Hs), axis=0))
    C, S = cos_sin_poly(Hs, degree)
    assert np.max(np.abs((C - 1j * S) - expm(-1j * Hs))) < 3e-16
    # squaring used for ||Hs|| > 0.65: (C - iS)^2 = (C^2 - S^2) - i (2 S C)
    U2 = (C @ C - S @ S) - 1j * (2.0 * S @ C)
    assert np.max(np.abs(U2 - expm(-2j * Hs))) < 1e-15


def test_real_symmetric_krylov_gradient_matches_the_oracle():
    p, eps = configs.c3_ensemble(n_delta=3, n_amp=2, NT=400)
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    K, N, L, NT = p.K, p.N, p.L, p.NT
    e = eps.reshape(L, NT)
    dts = np.diff(p.tlist)
    G = np.zeros((L, NT))
    for k in range(K):
        Hn = [(p.H0[k] + sum(e[l, n] * p.Hc[k, l] for l in range(L))).real for n in range(NT)]
        psi = p.psi0[k].copy()
        for n in range(NT):                                   # forward with the cos/sin propagators
            th = dts[n] * np.max(np.sum(np.abs(Hn[n]), axis=0))
            deg = 3 if th <= 2e-4 else 7 if th <= 3.5e-2 else 11 if th <= 0.23 else 15
            C, S = cos_sin_poly(dts[n] * Hn[n], deg)
            psi = (C - 1j * S) @ psi
        tau = np.vdot(p.tgt[k], psi)
        assert abs(tau - ref["tau"][k]) < 1e-13
        chi = tau * p.tgt[k] / K                              # J_T_ss
        rho = np.linalg.norm(chi)
        chi = chi / rho
        for n in range(NT - 1, -1, -1):                       # sym_step<N, M>
            dt = dts[n]
            th = dt * np.max(np.sum(np.abs(Hn[n]), axis=0))
            assert th <= TH[8]
            m = 2
            for j in range(2, 8):
                if th > TH[j]:
                    m = j + 1
            Hs = dt * Hn[n]
            w, ap = psi.copy(), psi.copy()
            E = [kappa(0, b) * w for b in range(m)]
            for a in range(1, m + 1):
                w = Hs @ w
                ap = ap + rot(w, a) / FACT[a]
                for b in range(m - a):
                    E[b] = E[b] + kappa(a, b) * rot(w, a)
            x, ac, IM = chi.copy(), chi.copy(), np.zeros((N, N))
            for b in range(m):
                z = rot(x, b)
                IM += np.outer(E[b].imag, z.real) - np.outer(E[b].real, z.imag)
                x = Hs @ x
                ac = ac + rot(x, b + 1) / FACT[b + 1]
            for l in range(L):
                G[l, n] += -2.0 * rho * dt * np.sum(p.Hc[k, l].real * IM)
            psi, chi = ap, ac
        assert np.max(np.abs(psi - p.psi0[k])) < 1e-12        # the forward state is carried back to t = 0
    Gr = ref["G"].reshape(L, NT)
    assert np.max(np.abs(G - Gr)) <= 1e-12 * np.max(np.abs(Gr))


@pytest.mark.parametrize("L", [1, 2, 3])
def test_operator_powers_from_symmetrised_products(L):
    """dense_dual_setup / dense_preform: H^2 = sum_{i<=j} c_i c_j P_ij, H^3 = sum_{i<=j<=k} c_i c_j c_k S_ijk, and the
    multi-term stage reproduces the Taylor recursion"""
    import itertools
    rng = np.random.default_rng(10 + L)
    N = 7
    Hs = [rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)) for _ in range(L + 1)]
    c = np.concatenate([[1.0], rng.standard_normal(L)])
    H = sum(ci * Hi for ci, Hi in zip(c, Hs))
    H2 = np.zeros((N, N), complex)
    H3 = np.zeros((N, N), complex)
    for i in range(L + 1):
        for j in range(i, L + 1):
            P = Hs[i] @ Hs[j] + (Hs[j] @ Hs[i] if i != j else 0)
            H2 += c[i] * c[j] * P
            for k in range(j, L + 1):
                S = sum(Hs[a] @ Hs[b] @ Hs[d] for a, b, d in set(itertools.permutations((i, j, k))))
                H3 += c[i] * c[j] * c[k] * S
    assert np.max(np.abs(H2 - H @ H)) < 1e-12 * np.max(np.abs(H @ H))
    assert np.max(np.abs(H3 - H @ H @ H)) < 1e-12 * np.max(np.abs(H @ H @ H))
    dt, m = 0.05, 11
    psi = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    terms = [psi]
    for j in range(1, m + 1):
        terms.append((-1j * dt / j) * (H @ terms[-1]))
    for NS in (2, 3):
        multi = [psi]
        j = 0
        while j < m:
            x = 1.0
            for q in range(min(NS, m - j)):
                x *= dt / (j + q + 1)
                Hq = (H, H2, H3)[q]
                multi.append(((-1j) ** (q + 1)) * x * (Hq @ multi[j]))
            j += NS
        assert len(multi) == m + 1
        assert max(np.max(np.abs(a - b)) for a, b in zip(multi, terms)) < 1e-14
