"""Deterministic synthetic inputs for the five BASELINE.json configs
 plus the reference's RNG-free test fixtures.

Every generator takes optional size overrides so that parity tests can run
reduced versions that the CPU oracle finishes in seconds."""
from __future__ import annotations

import numpy as np

from .problem import (GrapeProblem, SM, RE, SS, GRADGEN, TAYLOR, JA_NONE, JA_FLUENCE,
                      GB_NONE, GB_QUADFORM)


# -- pulse shapes (QuantumControl.Shapes) --------------------
def blackman(t, t0, T, a=0.16):
    t = np.asarray(t, dtype=np.float64)
    x = (t - t0) / (T - t0)
    v = 0.5 * (1.0 - a - np.cos(2 * np.pi * x) + a * np.cos(4 * np.pi * x))
    return np.where((t >= t0) & (t <= T), v, 0.0)


def flattop(t, T, t_rise, t0=0.0):
    t = np.asarray(t, dtype=np.float64)
    up = blackman(t, t0, t0 + 2 * t_rise)
    down = blackman(t, T - 2 * t_rise, T)
    v = np.ones_like(t)
    v = np.where(t < t0 + t_rise, up, v)
    v = np.where(t > T - t_rise, down, v)
    return np.where((t >= t0) & (t <= T), v, 0.0)


def midpoints(tlist):
    """Sampling points of `discretize_on_midpoints` (reference
    docs/src/background.md:55): interval mid-points, except first/last interval
    which are sampled at t_0 / t_NT."""
    tl = np.asarray(tlist, dtype=np.float64)
    m = 0.5 * (tl[:-1] + tl[1:])
    m[0], m[-1] = tl[0], tl[-1]
    return m


def discretize_on_midpoints(func, tlist):
    return np.asarray(func(midpoints(tlist)), dtype=np.float64) * np.ones(len(tlist) - 1)


# -- C1: README two-level |0> -> |1>  (reference README.md:40-59) ----------------
def c1_readme(NT=500, **kw):
    tlist = np.linspace(0.0, 5.0, NT + 1)
    H0 = np.diag([1.0, -1.0]).astype(np.complex128)
    H1 = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    psi0 = np.array([[1, 0]], dtype=np.complex128)
    tgt = np.array([[0, 1]], dtype=np.complex128)
    p = GrapeProblem(tlist, H0, H1[None], psi0, tgt, functional=SM, name="c1_readme_tls", **kw)
    return p, np.full(NT, 0.2)


# -- TLS fixture of test/test_tls_optimization.jl:20-36 --------------------------
def tls_fixture(NT=500, **kw):
    tlist = np.linspace(0.0, 5.0, NT + 1)
    H0 = -0.5 * np.diag([1.0, -1.0]).astype(np.complex128)
    H1 = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    psi0 = np.array([[1, 0]], dtype=np.complex128)
    tgt = np.array([[0, 1]], dtype=np.complex128)
    p = GrapeProblem(tlist, H0, H1[None], psi0, tgt, functional=SM, name="tls_fixture", **kw)
    eps = discretize_on_midpoints(lambda t: 0.2 * flattop(t, T=5.0, t_rise=0.3), tlist)
    return p, eps


# -- C2: 6-level transmon X gate, 4 basis trajectories ---------------------------
def c2_transmon(NT=2000, N=6, **kw):
    T = 20.0
    tlist = np.linspace(0.0, T, NT + 1)
    alpha = -2 * np.pi * 0.2
    j = np.arange(N)
    H0 = np.diag(alpha / 2 * j * (j - 1)).astype(np.complex128)
    a = np.diag(np.sqrt(np.arange(1, N)), 1).astype(np.complex128)
    Hx = (a + a.conj().T) / 2
    Hy = 1j * (a.conj().T - a) / 2
    s2 = 1 / np.sqrt(2)
    psi0 = np.zeros((4, N), dtype=np.complex128)
    psi0[0, 0] = 1
    psi0[1, 1] = 1
    psi0[2, 0], psi0[2, 1] = s2, s2
    psi0[3, 0], psi0[3, 1] = s2, 1j * s2
    X = np.eye(N, dtype=np.complex128)
    X[:2, :2] = [[0, 1], [1, 0]]
    tgt = psi0 @ X.T
    kw.setdefault("functional", SM)
    p = GrapeProblem(tlist, H0, np.stack([Hx, Hy]), psi0, tgt, name="c2_transmon_xgate", **kw)
    ex = discretize_on_midpoints(lambda t: (np.pi / T) * flattop(t, T=T, t_rise=2.0), tlist)
    ey = np.zeros(NT)
    return p, np.concatenate([ex, ey])


# -- C3: robust ensemble of 3-level Lambda systems -------------------------------
def c3_ensemble(n_delta=64, n_amp=64, NT=1000, functional=SS, **kw):
    """Lambda system of test/test_state_running_cost.jl:183-227 reduced to the two
    real controls; n_delta x n_amp grid of (detuning offset, amplitude scale)."""
    tlist = np.linspace(0.0, 5.0, NT + 1)
    dP = (10.0 - 0.0) - 9.5
    dS = (10.0 - 5.0) - 4.5
    H1P = 0.5 * np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0]], dtype=np.complex128)
    H1S = 0.5 * np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0]], dtype=np.complex128)
    deltas = np.linspace(-0.5, 0.5, n_delta) if n_delta > 1 else np.array([0.0])
    amps = np.linspace(0.9, 1.1, n_amp) if n_amp > 1 else np.array([1.0])
    K = n_delta * n_amp
    H0 = np.zeros((K, 3, 3), dtype=np.complex128)
    Hc = np.zeros((K, 2, 3, 3), dtype=np.complex128)
    k = 0
    for d in deltas:
        for a in amps:
            H0[k] = np.diag([0.0, dP + d, dP - dS])
            Hc[k, 0] = a * H1P
            Hc[k, 1] = a * H1S
            k += 1
    psi0 = np.zeros((K, 3), dtype=np.complex128)
    psi0[:, 0] = 1
    tgt = np.zeros((K, 3), dtype=np.complex128)
    tgt[:, 2] = 1
    p = GrapeProblem(tlist, H0, Hc, psi0, tgt, gen_of_traj=np.arange(K), functional=functional,
                     name=f"c3_ensemble_{K}", **kw)
    eP = discretize_on_midpoints(lambda t: blackman(t, 1.0, 5.0), tlist)
    eS = discretize_on_midpoints(lambda t: blackman(t, 0.0, 4.0), tlist)
    return p, np.concatenate([eP, eS])


# -- C4 / C5: dense synthetic Hamiltonians ---------------------------------------
def _gue(N, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    H = (A + A.conj().T) / 2
    return H / np.max(np.abs(np.linalg.eigvalsh(H)))


def _dense(N, K, NT, seeds, name, **kw):
    H0, H1, H2 = (_gue(N, s) for s in seeds[:3])
    Q, _ = np.linalg.qr(_gue(N, seeds[3]))
    tlist = 0.25 * np.arange(NT + 1)
    psi0 = np.eye(N, dtype=np.complex128)[:K]
    tgt = np.ascontiguousarray(Q[:, :K].T)
    kw.setdefault("functional", SM)
    p = GrapeProblem(tlist, H0, np.stack([H1, H2]), psi0, tgt, name=name, **kw)
    n = np.arange(NT)
    eps = np.concatenate([0.5 * np.sin(np.pi * (l + 2) * (n + 0.5) / NT) for l in range(2)])
    # H_n = H0 + e1 H1 + e2 H2 with |e| <= 0.5: ||H_n dt||_2 <= 0.25 * 2 = 0.5
    return p, eps


def c4_dense450(N=450, K=16, NT=5000, **kw):
    return _dense(N, K, NT, (1001, 1002, 1003, 1004), f"c4_dense{N}", **kw)


def c5_dense1024(N=1024, K=64, NT=1000, **kw):
    D = np.zeros((N, N), dtype=np.complex128)
    idx = np.arange(N // 2, N)
    D[idx, idx] = 1.0
    kw.setdefault("ja_kind", JA_FLUENCE)
    kw.setdefault("lambda_a", 1e-2)
    kw.setdefault("gb_kind", GB_QUADFORM)
    kw.setdefault("lambda_b", 0.1)
    kw.setdefault("gb_D", D)
    return _dense(N, K, NT, (2001, 2002, 2003, 2004), f"c5_dense{N}", **kw)


def random_problem(K=3, N=3, L=2, NT=12, G=None, seed=0, hermitian=True, uniform=False,
                   shaped=False, real=False, **kw):
    """Randomised small problem (non-uniform grid, optionally non-Hermitian
    generators, cf. reference test/test_taylor_grad.jl:17-20; `real`: real-symmetric
    generators, the rotating-frame / real-control case)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = K if G is None else G

    def rmat():
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        if real:
            A = A.real + 0j
        return (A + A.conj().T) / 2 if (hermitian or real) else A

    H0 = np.stack([rmat() for _ in range(G)])
    Hc = np.stack([np.stack([rmat() for _ in range(L)]) for _ in range(G)])
    dts = np.full(NT, 0.05) if uniform else 0.02 + 0.06 * rng.random(NT)
    tlist = np.concatenate([[0.0], np.cumsum(dts)])

    def rstate():
        v = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    psi0, tgt = rstate(), rstate()
    gen = rng.integers(0, G, size=K) if G not in (1, K) else None
    shape = 0.5 + rng.random((L, NT)) if shaped else None
    weights = kw.pop("weights", None)
    p = GrapeProblem(tlist, H0, Hc, psi0, tgt, gen_of_traj=gen, shape=shape, weights=weights,
                     name=f"random_K{K}_N{N}_L{L}", **kw)
    eps = rng.standard_normal(L * NT)
    return p, eps


def lindblad_tls(NT=200, T=5.0, gamma=0.05, K=2, **kw):
    """Open two-level system as Liouville-space trajectories (reference docs/src/background.md:46, 240-242:
    density matrices are propagated as vectors under a Liouvillian): rho' = -i[H, rho] + gamma (s rho s^+ -
    {s^+ s, rho}/2), H = -sigma_z/2 + eps(t) sigma_x, row-major vec(rho) of length N = 4.  The propagator of a step
    is exp(-i G dt) with the NON-HERMITIAN generator G = i L (L the Liouvillian), which is linear in the control:
    G = (H (x) 1 - 1 (x) H^T) + i D.  Trajectories: rho(0) = |0><0| -> |1><1| and the maximally mixed state -> itself
    ... K = 2; `J_T_re` with tau_k = <<rho_tgt | rho(T)>> = Tr(rho_tgt rho(T))."""
    tlist = np.linspace(0.0, T, NT + 1)
    sz = np.diag([1.0, -1.0]).astype(np.complex128)
    sx = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    sm = np.array([[0, 1], [0, 0]], dtype=np.complex128)       # |0><1|: decay 1 -> 0
    I2 = np.eye(2, dtype=np.complex128)

    def comm(H):                                               # vec_r(H rho - rho H) = (H (x) 1 - 1 (x) H^T) vec_r(rho)
        return np.kron(H, I2) - np.kron(I2, H.T)

    n_op = sm.conj().T @ sm
    D = gamma * (np.kron(sm, sm.conj()) - 0.5 * np.kron(n_op, I2) - 0.5 * np.kron(I2, n_op.T))
    G0 = comm(-0.5 * sz) + 1j * D
    G1 = comm(sx)
    rho0 = [np.diag([1.0, 0.0]), 0.5 * np.eye(2)][:K]
    rhoT = [np.diag([0.0, 1.0]), 0.5 * np.eye(2)][:K]
    psi0 = np.array([r.reshape(-1) for r in rho0], dtype=np.complex128)
    tgt = np.array([r.reshape(-1) for r in rhoT], dtype=np.complex128)
    kw.setdefault("functional", RE)
    p = GrapeProblem(tlist, G0, G1[None], psi0, tgt, name="lindblad_tls", **kw)
    eps = discretize_on_midpoints(lambda t: 0.3 * flattop(t, T=T, t_rise=0.3), tlist)
    return p, eps


CONFIGS = {
    "c1": c1_readme,
    "c2": c2_transmon,
    "c3": c3_ensemble,
    "c4": c4_dense450,
    "c5": c5_dense1024,
}
