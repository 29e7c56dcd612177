"""Spectral CPU oracle for shared-generator dense problems -- TEST INFRASTRUCTURE ONLY.

`oracle/grape_oracle.py` restates the reference literally: per (trajectory, step) one
dense Pade `expm` of the N x N generator forward (optimize.jl:732) and one of the
N(L+1) x N(L+1) GradGenerator block matrix backward (optimize.jl:881).  At the full
sizes of BASELINE configs[3] / [4] (N = 450 x 5000 steps, N = 1024 x 1000 steps) that
is weeks of CPU time.  This module computes THE SAME QUANTITIES (same index
conventions, same trapezoid weights, same chi boundary / normalisation / xi terms,
cited line by line below) with an algorithm that shares nothing with the CUDA path's
Taylor/Krylov series:

  * the generator of a step is shared by all trajectories and Hermitian, so
    H_n = V diag(lam) V^dagger (LAPACK zheevd) once per step,
    U_n = V diag(exp(-i lam dt)) V^dagger                                 (ExpProp prop_step!)
  * the GradGenerator step  [chi'_1..chi'_L; chi] <- exp(-i G (-dt)) [0..0; chi]
    (docs/src/background.md:447-494) is   chi <- U_n^dagger chi,
    chi'_l = (dU_n^dagger / d eps_l) chi,  and the Frechet derivative of the exponential
    of a normal matrix is exact in its eigenbasis (Daleckii-Krein):
      dU/d eps_l = V [ (V^dagger (-i dt mu_l) V) o Phi ] V^dagger,
      Phi_ij = (e^{x_i} - e^{x_j})/(x_i - x_j) = e^{(x_i+x_j)/2} sinc(dt (lam_i-lam_j)/2),  x = -i dt lam
    (the sinc form has no cancellation for near-degenerate pairs).

It is validated against `grape_oracle.evaluate_gradient` (both gradient methods) on
reduced sizes in tests/test_dense_oracle.py and produces the full-size golden vectors
tests/golden/dense_full/c4_dense450_full.npz / c5_dense1024_full.npz
(tests/golden/make_golden_dense.py).  Never imported by the product."""
from __future__ import annotations

import numpy as np

SM, RE, SS = 0, 1, 2


def _amps(p, eps, n):
    a = np.array([eps[l * p.NT + n] for l in range(p.L)])
    s = np.ones(p.L) if p.shape is None else p.shape[:, n]
    return a * s, s


def evaluate_gradient(p, pulsevals, keep_eig=True, progress=None):
    """Shared (G == 1) Hermitian generator. Returns dict(J, J_parts, tau, G, grad_J_Tb, grad_J_a,
    final_states, chi_norms).  `p`: GrapeProblem / OracleProblem attribute names."""
    assert p.G == 1, "spectral oracle: one shared generator"
    eps = np.asarray(pulsevals, dtype=np.float64)
    K, N, L, NT, tl = p.K, p.N, p.L, p.NT, np.asarray(p.tlist)
    H0, Hc = p.H0[0], p.Hc[0]
    assert np.allclose(H0, H0.conj().T) and all(np.allclose(h, h.conj().T) for h in Hc), "Hermitian generators only"
    w = np.ones(K) if getattr(p, "weights", None) is None else np.asarray(p.weights)
    Kg = float(getattr(p, "K_global", K) or K)
    use_gb = p.gb_kind != 0
    D = None
    if use_gb:
        D = np.asarray(p.gb_D)
        assert D.shape[0] == 1, "spectral oracle: one shared D"
        D = D[0]

    def g_b(Psi):   # [N,K] -> [K]   g_b = <Psi|D|Psi>  (test/test_state_running_cost.jl:17-30)
        return np.real(np.einsum("ik,ik->k", Psi.conj(), D @ Psi))

    # ---- evaluate_functional (optimize.jl:696-768)
    store = np.empty((NT + 1, N, K), dtype=np.complex128)           # fw_storage, column n = Psi(t_n)  :723, :738
    Psi = np.ascontiguousarray(p.psi0.T.copy())
    store[0] = Psi
    Jb = np.zeros(K)
    if use_gb:
        Jb += g_b(Psi) * ((tl[1] - tl[0]) / 2)                      # :727-730
    eigs = [None] * NT
    for n in range(NT):                                             # :731
        dt = tl[n + 1] - tl[n]
        a, _ = _amps(p, eps, n)
        H = H0 + np.tensordot(a, Hc, axes=1)
        lam, V = np.linalg.eigh(H)
        if keep_eig:
            eigs[n] = (lam, V)
        Psi = V @ (np.exp(-1j * lam * dt)[:, None] * (V.conj().T @ Psi))   # prop_step!  :732
        store[n + 1] = Psi
        if use_gb:                                                  # :739-750
            wt = 0.5 * (tl[n + 2] - tl[n]) if n + 1 < NT else (tl[-1] - tl[-2]) / 2
            Jb += g_b(Psi) * wt
        if progress and (n % progress == 0):
            print(f"  forward {n}/{NT}", flush=True)
    tgt = np.ascontiguousarray(p.tgt.T)
    tau = np.einsum("ik,ik->k", tgt.conj(), Psi)                    # :753
    J_parts = np.zeros(3)
    if p.functional == SM:
        J_parts[0] = 1.0 - abs(np.sum(w * tau) / Kg) ** 2
        c = w * np.sum(w * tau) / Kg ** 2
    elif p.functional == RE:
        J_parts[0] = 1.0 - np.real(np.sum(w * tau)) / Kg
        c = w / (2.0 * Kg) + 0j
    else:
        J_parts[0] = 1.0 - np.sum(w * np.abs(tau) ** 2) / Kg
        c = w * tau / Kg
    dts = np.diff(tl)
    e2 = eps.reshape(L, NT)
    grad_J_a = np.zeros(L * NT)
    if p.ja_kind:
        J_parts[1] = p.lambda_a * float(np.sum(e2 * e2 * dts[None, :]))     # :761-763
        grad_J_a = (2.0 * e2 * dts[None, :]).reshape(-1)
    if use_gb:
        J_parts[2] = p.lambda_b * float(np.sum(Jb))                 # :764-766

    # ---- evaluate_gradient! (optimize.jl:824-1014)
    chi = tgt * c[None, :]                                          # :845-855
    use_xi = use_gb and p.lambda_b != 0.0
    if use_xi:                                                      # :856-866
        chi = chi + (p.lambda_b * (tl[-1] - tl[-2]) / 2) * (-(D @ Psi))
    rho = np.linalg.norm(chi, axis=0)                               # :867
    if np.any(rho < p.chi_min_norm):
        raise RuntimeError("chi norm below chi_min_norm")           # :1021-1025
    chi = chi / rho[None, :]
    gT = np.zeros((L, NT))
    for n in range(NT - 1, -1, -1):                                 # :880
        dt = tl[n + 1] - tl[n]
        a, s = _amps(p, eps, n)
        if eigs[n] is not None:
            lam, V = eigs[n]
            eigs[n] = None
        else:
            lam, V = np.linalg.eigh(H0 + np.tensordot(a, Hc, axes=1))
        Vh = V.conj().T
        psi_prev = store[n]                                         # fw_storage[k][:, n] = Psi(t_{n-1}) 1-based  :888-892
        A = Vh @ psi_prev                                           # [N,K]
        C = Vh @ chi
        x = -1j * dt * lam
        th = 0.5 * dt * (lam[:, None] - lam[None, :])
        Phi = np.exp(0.5 * (x[:, None] + x[None, :])) * np.sinc(th / np.pi)
        # sum_k rho_k <chi'_{kl}|Psi_k> = sum_k rho_k chi_k^dagger (dU/d eps_l) Psi_k
        #   = -i dt s_l sum_ij (V^dagger mu_l V)_ij Phi_ij sum_k rho_k conj(C_ik) A_jk = -i dt s_l sum_pq mu_l[p,q] R[p,q]
        W = Phi * ((C.conj() * rho[None, :]) @ A.T)
        R = V.conj() @ W @ V.T
        for l in range(L):
            gT[l, n] = -2.0 * np.real(-1j * dt * s[l] * np.sum(Hc[l] * R))      # :893-895, :574-584
        chi = V @ (np.exp(1j * lam * dt)[:, None] * C)              # chi <- U_n^dagger chi  (:881, state block)
        if use_xi and n > 0:                                        # :897-908
            wt = 0.5 * (tl[n + 1] - tl[n - 1])
            chi = chi + (p.lambda_b * wt) * (-(D @ psi_prev)) / rho[None, :]
        if progress and (n % progress == 0):
            print(f"  backward {n}/{NT}", flush=True)
    grad_J_Tb = gT.reshape(-1)
    G = grad_J_Tb + (p.lambda_a * grad_J_a if p.ja_kind else 0.0)   # :1003-1011
    return dict(J=float(np.sum(J_parts)), J_parts=J_parts, tau=tau, G=G, grad_J_Tb=grad_J_Tb, grad_J_a=grad_J_a,
                final_states=np.ascontiguousarray(Psi.T), chi_norms=rho)
