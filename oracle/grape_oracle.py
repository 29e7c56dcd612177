"""CPU oracle for the GRAPE gradient hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain NumPy/SciPy (complex128) restatement of the algorithm that
GRAPE.jl executes for `prop_method=ExpProp` -- it is *not* part of the product.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it.  The product path
(`grape.jl_b200`) never imports anything from `oracle/`.

PARITY STATUS: **parity unpinned element-wise.**  The reference
(/root/reference) stores no golden J_T / gradient / pulse vectors, Julia is not
installed in this image, and the arithmetic lives in un-vendored, un-pinned
dependencies (QuantumPropagators.jl `ExpProp`, QuantumGradientGenerators.jl
`GradGenerator`, QuantumControl.jl functionals, Julia `LinearAlgebra.exp`;
Project.toml:21-30, no Manifest).  The oracle is therefore pinned by
  (1) every RNG-free known-answer / threshold / identity test the reference
      holds for this path (tests/test_oracle_known_answers.py):
      test/test_tls_optimization.jl:169-170, :229, :260;
      test/test_readme_example.jl:37-38; test/test_taylor_grad.jl:33-69;
      test/test_state_running_cost.jl:41-48;
  (2) central finite differences of its own functional;
  (3) `:taylor` == `:gradgen`;
  (4) the analytic Rabi answer for the README problem.

Reference lines each function follows are cited in its docstring
(`optimize.jl` = /root/reference/src/optimize.jl, etc.).

Conventions:
  K trajectories, N levels, L controls, NT = len(tlist)-1 intervals.
  pulsevals[(l)*NT + n]   (0-based l, n)  -- blocked by control, workspace.jl:159-162.
  H_{k,n} = H0[g] + sum_l shape[l,n]*eps[l,n]*Hc[g][l],  g = gen_of_traj[k].
  storage[k][:, n] = Psi_k(t_n), n = 0..NT                (optimize.jl:723, 738)
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

SM, RE, SS, HOST = 0, 1, 2, 3          # functional kinds
GRADGEN, TAYLOR = 0, 1                 # gradient methods
JA_NONE, JA_FLUENCE = 0, 1
GB_NONE, GB_QUADFORM = 0, 1


class OracleProblem:
    """Plain container; every field is a NumPy array or scalar."""

    def __init__(self, tlist, H0, Hc, psi0, tgt, gen_of_traj=None, shape=None,
                 weights=None, functional=SM, gradient_method=GRADGEN,
                 ja_kind=JA_NONE, lambda_a=1.0, gb_kind=GB_NONE, lambda_b=1.0,
                 gb_D=None, chi_min_norm=1e-100, taylor_max_order=100,
                 taylor_tolerance=1e-16, taylor_check_convergence=True,
                 K_global=None):
        self.tlist = np.asarray(tlist, dtype=np.float64)
        H0 = np.asarray(H0, dtype=np.complex128)
        if H0.ndim == 2:
            H0 = H0[None]
        Hc = np.asarray(Hc, dtype=np.complex128)
        if Hc.ndim == 3:
            Hc = Hc[None]
        self.H0, self.Hc = H0, Hc                      # [G,N,N], [G,L,N,N]
        self.psi0 = np.asarray(psi0, dtype=np.complex128)   # [K,N]
        self.tgt = np.asarray(tgt, dtype=np.complex128)     # [K,N]
        self.K, self.N = self.psi0.shape
        self.G, self.L = Hc.shape[0], Hc.shape[1]
        self.NT = len(self.tlist) - 1
        if gen_of_traj is None:
            gen_of_traj = np.arange(self.K) if self.G == self.K and self.G > 1 \
                else np.zeros(self.K, dtype=np.int64)
        self.gen_of_traj = np.asarray(gen_of_traj, dtype=np.int64)
        self.shape = None if shape is None else np.asarray(shape, dtype=np.float64).reshape(self.L, self.NT)
        self.weights = np.ones(self.K) if weights is None else np.asarray(weights, dtype=np.float64)
        self.functional = functional
        self.gradient_method = gradient_method
        self.ja_kind, self.lambda_a = ja_kind, float(lambda_a)
        self.gb_kind, self.lambda_b = gb_kind, float(lambda_b)
        if gb_D is not None:
            gb_D = np.asarray(gb_D, dtype=np.complex128)
            if gb_D.ndim == 2:
                gb_D = gb_D[None]
        self.gb_D = gb_D                                # [1 or K, N, N]
        self.chi_min_norm = chi_min_norm
        self.taylor_max_order = taylor_max_order
        self.taylor_tolerance = taylor_tolerance
        self.taylor_check_convergence = taylor_check_convergence
        # for a trajectory shard: the functionals are normalised by the GLOBAL K
        self.K_global = self.K if K_global is None else int(K_global)
        # amplitude mode (non-linear controls, workspace.jl:283-285 / optimize.jl:946-951): when set, `pulsevals`
        # holds the amplitudes a_{l,n} themselves and dampl[l,n] = d a_{l,n} / d eps scales the control derivative
        self.dampl = None


def from_problem(p, **overrides):
    """Build an OracleProblem from any object with the same attribute names."""
    names = ("tlist H0 Hc psi0 tgt gen_of_traj shape weights functional gradient_method "
             "ja_kind lambda_a gb_kind lambda_b gb_D chi_min_norm taylor_max_order "
             "taylor_tolerance taylor_check_convergence K_global").split()
    kw = {n: getattr(p, n) for n in names if hasattr(p, n)}
    kw.update(overrides)
    return OracleProblem(**kw)


# ----------------------------------------------------------------------------
# generator evaluation  (QuantumPropagators `evaluate(generator, tlist, n; vals_dict)`,
# called inside prop_step! at optimize.jl:732, 881 and explicitly at :942-951)
# ----------------------------------------------------------------------------
def amplitude(p, pulsevals, l, n):
    eps = pulsevals[l * p.NT + n]
    if getattr(p, "dampl", None) is not None:
        return eps
    return eps if p.shape is None else p.shape[l, n] * eps


def generator(p, pulsevals, k, n):
    g = p.gen_of_traj[k]
    H = p.H0[g].copy()
    for l in range(p.L):
        H += amplitude(p, pulsevals, l, n) * p.Hc[g, l]
    return H


def control_deriv(p, k, l, n):
    """mu_{k,l,n} = dH/d eps_{l,n}  (get_control_derivs, workspace.jl:283-285)."""
    g = p.gen_of_traj[k]
    if getattr(p, "dampl", None) is not None:      # evaluate(mu) of a non-linear amplitude; 0: `isnothing(mu)` branch
        return p.dampl[l, n] * p.Hc[g, l]
    s = 1.0 if p.shape is None else p.shape[l, n]
    return s * p.Hc[g, l]


# ----------------------------------------------------------------------------
# functionals (QuantumControl.Functionals; formulas restated in
# docs/src/tutorial.md:326-367, 399-405 and SURVEY.md section 8a)
# ----------------------------------------------------------------------------
def J_T_from_tau(kind, tau, weights, K_global):
    if kind == SM:
        f = np.sum(weights * tau) / K_global
        return 1.0 - abs(f) ** 2
    if kind == RE:
        f = np.sum(weights * tau) / K_global
        return 1.0 - f.real
    if kind == SS:
        return 1.0 - np.sum(weights * np.abs(tau) ** 2) / K_global
    raise ValueError("HOST functional has no built-in J_T")


def chi_coeff_from_tau(kind, tau, weights, K_global, sigma=None):
    """chi_k(T) = c_k * tgt_k   with chi_k = -dJ_T/d<Psi_k|  (make_chi, workspace.jl:306-308).

    `sigma` = global sum_j w_j tau_j (pass it in when `tau` is only a shard)."""
    if kind == SM:
        if sigma is None:
            sigma = np.sum(weights * tau)
        return weights * sigma / K_global ** 2
    if kind == RE:
        return weights / (2.0 * K_global) + 0j
    if kind == SS:
        return weights * tau / K_global
    raise ValueError("HOST functional has no built-in chi")


def g_b_value(p, k, psi):
    """g_b(Psi) = <Psi|D|Psi>  (test/test_state_running_cost.jl:17-30)."""
    D = p.gb_D[0 if p.gb_D.shape[0] == 1 else k]
    return float(np.real(np.vdot(psi, D @ psi)))


def xi_value(p, k, psi):
    """xi = -d g_b / d<Psi| = -D Psi  (docs/src/background.md:613-778; make_xi, workspace.jl:312-324)."""
    D = p.gb_D[0 if p.gb_D.shape[0] == 1 else k]
    return -(D @ psi)


def J_a_value(p, pulsevals):
    """J_a_fluence = sum eps^2 dt_n  (QuantumControl.Functionals.J_a_fluence)."""
    if p.ja_kind == JA_NONE:
        return 0.0
    dt = np.diff(p.tlist)
    e = np.asarray(pulsevals).reshape(p.L, p.NT)
    return float(np.sum(e * e * dt[None, :]))


def grad_J_a_value(p, pulsevals):
    if p.ja_kind == JA_NONE:
        return np.zeros(p.L * p.NT)
    dt = np.diff(p.tlist)
    e = np.asarray(pulsevals).reshape(p.L, p.NT)
    return (2.0 * e * dt[None, :]).reshape(-1)


# ----------------------------------------------------------------------------
# evaluate_functional   (optimize.jl:696-768)
# ----------------------------------------------------------------------------
def evaluate_functional(p, pulsevals, want_storage=True, sigma_reduce=None, jt_host=None):
    """Forward sweep. Returns dict(J, J_parts[3], tau[K], storage[K,N,NT+1], J_b_trajectory[K]).

    `sigma_reduce(partial)->global` lets a trajectory shard obtain the global
    J_T ingredients (used by the world_size-2 tests); default: identity."""
    pulsevals = np.asarray(pulsevals, dtype=np.float64)
    K, N, NT, tl = p.K, p.N, p.NT, p.tlist
    storage = np.zeros((K, N, NT + 1), dtype=np.complex128)
    tau = np.zeros(K, dtype=np.complex128)
    J_b_traj = np.zeros(K)
    for k in range(K):
        psi = p.psi0[k].copy()
        storage[k, :, 0] = psi                                       # :723
        if p.gb_kind != GB_NONE:
            J_b_traj[k] = g_b_value(p, k, psi) * ((tl[1] - tl[0]) / 2)   # :727-730
        for n in range(NT):                                          # :731
            dt = tl[n + 1] - tl[n]
            U = expm(-1j * generator(p, pulsevals, k, n) * dt)        # ExpProp prop_step!  :732
            psi = U @ psi
            storage[k, :, n + 1] = psi                               # :738
            if p.gb_kind != GB_NONE:                                 # :739-750
                n_tl = n + 1
                if n_tl < NT:
                    w = 0.5 * (tl[n_tl + 1] - tl[n_tl - 1])
                else:
                    w = (tl[-1] - tl[-2]) / 2
                J_b_traj[k] += g_b_value(p, k, psi) * w
        tau[k] = np.vdot(p.tgt[k], psi)                              # :753
    J_parts = np.zeros(3)
    J_parts[0] = _J_T(p, tau, sigma_reduce, storage[:, :, -1], jt_host)   # :755-760
    if p.ja_kind != JA_NONE:
        J_parts[1] = p.lambda_a * J_a_value(p, pulsevals)            # :761-763
    if p.gb_kind != GB_NONE:
        J_parts[2] = p.lambda_b * np.sum(J_b_traj)                   # :764-766
    return dict(J=float(np.sum(J_parts)), J_parts=J_parts, tau=tau,
                storage=storage if want_storage else None, J_b_trajectory=J_b_traj,
                final_states=storage[:, :, -1].copy())


def _J_T(p, tau, sigma_reduce, final_states, jt_host):
    if p.functional == HOST:
        return float(jt_host(final_states))
    if sigma_reduce is None:
        return J_T_from_tau(p.functional, tau, p.weights, p.K_global)
    # sharded: reduce the additive ingredient, then finish
    if p.functional in (SM, RE):
        sigma = sigma_reduce(np.sum(p.weights * tau))
        f = sigma / p.K_global
        return 1.0 - abs(f) ** 2 if p.functional == SM else 1.0 - f.real
    s = sigma_reduce(np.sum(p.weights * np.abs(tau) ** 2) + 0j)
    return 1.0 - s.real / p.K_global


# ----------------------------------------------------------------------------
# taylor_grad_step!   (optimize.jl:604-653)
# ----------------------------------------------------------------------------
def taylor_grad_step(psi, H, mu, dt, check_convergence=True, max_order=100, tolerance=1e-16):
    """(d/d eps) exp(-i H dt) psi  with  mu = dH/d eps ; Kuprov & Rodgers recursion."""
    phi_prev = mu @ psi                  # Phi_1                               :619
    Hn1_psi = H @ psi                    # H^{n-1} psi, n=2                    :620
    alpha = -1j * dt                     # :621
    out = alpha * phi_prev               # :622
    r = 0.0
    for n in range(2, max_order + 1):    # :626
        phi = H @ phi_prev + mu @ Hn1_psi        # :628-629
        alpha = alpha * (-1j * dt / n)           # :631
        out = out + alpha * phi                  # :632
        if check_convergence:
            r = abs(alpha * np.linalg.norm(phi))  # :634
            if r < tolerance:
                return out
        Hn1_psi = H @ Hn1_psi                    # :639
        phi_prev = phi
    if check_convergence and max_order > 1:
        raise RuntimeError(
            f"taylor_grad_step! did not converge within {max_order} iterations. Residual term r={r}.")
    return out


# ----------------------------------------------------------------------------
# evaluate_gradient!   (optimize.jl:824-1014)
# ----------------------------------------------------------------------------
def gradgen_matrix(A, Bs):
    """Dense block upper-triangular GradGenerator matrix, docs/src/background.md:467-477."""
    N, L = A.shape[0], len(Bs)
    Gm = np.zeros((N * (L + 1), N * (L + 1)), dtype=np.complex128)
    for l in range(L + 1):
        Gm[l * N:(l + 1) * N, l * N:(l + 1) * N] = A
    for l in range(L):
        Gm[l * N:(l + 1) * N, L * N:(L + 1) * N] = Bs[l]
    return Gm


def evaluate_gradient(p, pulsevals, sigma_reduce=None, jt_host=None, chi_host=None):
    """Returns dict(J, G, J_parts, tau, grad_J_Tb, grad_J_a, tau_grads[K,NT,L],
    chi_states[K,N], chi_norms[K], storage)."""
    pulsevals = np.asarray(pulsevals, dtype=np.float64)
    K, N, NT, L, tl = p.K, p.N, p.NT, p.L, p.tlist
    fw = evaluate_functional(p, pulsevals, True, sigma_reduce, jt_host)    # :842-843
    storage, tau = fw["storage"], fw["tau"]
    use_xi = p.gb_kind != GB_NONE and p.lambda_b != 0.0                     # :831-836

    # chi_k(T)                                                              :845-855
    if p.functional == HOST:
        chi = np.array(chi_host(storage[:, :, -1]), dtype=np.complex128)
    else:
        sigma = None
        if sigma_reduce is not None and p.functional == SM:
            sigma = sigma_reduce(np.sum(p.weights * tau))
        c = chi_coeff_from_tau(p.functional, tau, p.weights, p.K_global, sigma)
        chi = c[:, None] * p.tgt
    if use_xi:                                                              # :856-866
        dtl = tl[-1] - tl[-2]
        for k in range(K):
            chi[k] = chi[k] + (p.lambda_b * dtl / 2) * xi_value(p, k, storage[k, :, -1])
    rho = np.linalg.norm(chi, axis=1)                                       # :867
    for k in range(K):                                                      # :1017-1038
        if rho[k] < p.chi_min_norm:
            raise RuntimeError(
                f"The χ state with index {k + 1} has norm {rho[k]} < {p.chi_min_norm} (chi_min_norm)")
        chi[k] = chi[k] / rho[k]
    chi_states = chi.copy()

    tau_grads = np.zeros((K, NT, L), dtype=np.complex128)
    for k in range(K):
        x = chi[k].copy()
        for n in range(NT - 1, -1, -1):                                     # :880 / :924
            dt = tl[n + 1] - tl[n]
            A = generator(p, pulsevals, k, n).conj().T       # adjoint generator, workspace.jl:153
            Bs = [control_deriv(p, k, l, n).conj().T for l in range(L)]
            psi_prev = storage[k, :, n]                      # = Psi(t_{n-1}) 1-based  :888-892
            if p.gradient_method == GRADGEN:
                v = np.zeros(N * (L + 1), dtype=np.complex128)
                v[L * N:] = x                                # GradVector(chi, L)  :878
                v = expm(-1j * gradgen_matrix(A, Bs) * (-dt)) @ v    # backward prop_step!  :881
                for l in range(L):
                    tau_grads[k, n, l] = rho[k] * np.vdot(v[l * N:(l + 1) * N], psi_prev)  # :893-895
                x = v[L * N:]                                # resetgradvec!  :896
            else:
                for l in range(L):                           # :946-970
                    xt = taylor_grad_step(x, A, Bs[l], -dt, p.taylor_check_convergence,
                                          p.taylor_max_order, p.taylor_tolerance)
                    tau_grads[k, n, l] = rho[k] * np.vdot(xt, psi_prev)
                x = expm(-1j * A * (-dt)) @ x                # bw prop_step!  :972
            if use_xi and n > 0:                             # :897-908 / :979-992  (n>1 1-based)
                w = 0.5 * (tl[n + 1] - tl[n - 1])
                x = x + (p.lambda_b * w / rho[k]) * xi_value(p, k, psi_prev)

    grad_J_Tb = np.zeros(L * NT)                                            # :574-584
    for l in range(L):
        grad_J_Tb[l * NT:(l + 1) * NT] = -2.0 * np.real(np.sum(tau_grads[:, :, l], axis=0))
    if sigma_reduce is not None:
        grad_J_Tb = sigma_reduce(grad_J_Tb)
    G = grad_J_Tb.copy()                                                    # :1003
    grad_J_a = grad_J_a_value(p, pulsevals)
    if p.ja_kind != JA_NONE:                                                # :1004-1011
        G += p.lambda_a * grad_J_a
    out = dict(fw)
    out.update(G=G, grad_J_Tb=grad_J_Tb, grad_J_a=grad_J_a, tau_grads=tau_grads,
               chi_states=chi_states, chi_norms=rho)
    return out


def finite_difference_gradient(p, pulsevals, idx, h=1e-6, **kw):
    """Central differences of the oracle's own J (gate G1 of SURVEY.md 7.1a)."""
    pulsevals = np.asarray(pulsevals, dtype=np.float64)
    out = np.zeros(len(idx))
    for j, i in enumerate(idx):
        xp, xm = pulsevals.copy(), pulsevals.copy()
        xp[i] += h
        xm[i] -= h
        out[j] = (evaluate_functional(p, xp, False, **kw)["J"]
                  - evaluate_functional(p, xm, False, **kw)["J"]) / (2 * h)
    return out
