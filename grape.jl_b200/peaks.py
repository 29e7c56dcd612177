"""FP64 roofline denominators measured on the GPU the bench runs on (csrc/peaks.cu)
and the model of the FP64 work the small-N kernels execute per unit."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgrape_peaks.so")


def measure(device=0):
    lib = C.CDLL(_SO)
    lib.gb_peak_dfma_tflops.restype = C.c_double
    lib.gb_peak_dmma_tflops.restype = C.c_double
    return dict(dfma_tflops=lib.gb_peak_dfma_tflops(int(device)),
                dmma_tflops=lib.gb_peak_dmma_tflops(int(device)))


def _exp_plan(nrm):
    """mirror of exp_plan() in csrc/common.cuh -> (complex matmuls, squarings)"""
    deg = np.where(nrm <= 2e-4, 3, np.where(nrm <= 3.5e-2, 7, np.where(nrm <= 0.23, 11, 15)))
    s = np.where(nrm > 0.65, np.ceil(np.log2(np.maximum(nrm, 1e-300) / 0.65)), 0).astype(int)
    return deg, np.maximum(s, 0)


def _vec_terms(nrm):
    """mirror of vec_plan(): Taylor terms m and sub-steps 2^s"""
    s = np.where(nrm > 1.0, np.ceil(np.log2(np.maximum(nrm, 1e-300))), 0).astype(int)
    th = nrm / (2.0 ** s)
    m = np.ones_like(th, dtype=int)
    t = th.copy()
    for j in range(2, 41):
        need = t > 2e-17
        m = np.where(need, j, m)
        t = np.where(need, t * th / j, t)
    return np.maximum(m, 2), s


def _plans(p, eps, sample=64):
    N, L, NT, K, G = p.N, p.L, p.NT, p.K, p.G
    gs = np.unique(np.linspace(0, G - 1, min(G, sample)).astype(int))
    e = np.asarray(eps).reshape(L, NT)
    if p.shape is not None:
        e = e * p.shape
    dt = np.diff(p.tlist)
    H = p.H0[gs][:, None] + np.einsum("ln,glij->gnij", e, p.Hc[gs])
    nrm = np.max(np.sum(np.abs(H.real) + np.abs(H.imag), axis=2), axis=2) * dt[None, :]
    return nrm


def econ_on():
    """the kernels' series: economised (Chebyshev-cut) polynomial unless GRAPE_B200_ECON=0 (csrc/econ.cuh)"""
    return os.environ.get("GRAPE_B200_ECON", "1") != "0"


_ECON_THETA = None


def econ_theta():
    """theta[m], m = 0..20: the radius the economised polynomial of degree m serves (the library's own table)"""
    global _ECON_THETA
    if _ECON_THETA is None:
        from . import _lib
        lib = _lib.load()
        th = np.zeros(21)
        for m in range(2, 21):
            t = C.c_double()
            g = (C.c_double * (m + 1))()
            lib.grape_b200_econ_table(m, C.byref(t), g)
            th[m] = t.value
        _ECON_THETA = th
    return _ECON_THETA


def _order_for(theta, radii):
    """smallest index m >= 2 with theta <= radii[m] (len(radii) - 1 if none)"""
    m = np.full(theta.shape, len(radii) - 1, dtype=int)
    for j in range(len(radii) - 2, 1, -1):
        m = np.where(theta <= radii[j], j, m)
    return m


def _fact(n):
    return float(np.prod(np.arange(1, n + 1, dtype=float))) if n > 0 else 1.0


SYM_CLS_DEG = (3, 5, 6, 8, 12, 16)   # csrc/small_sym.cuh


def _sym_flops_per_unit(p, eps, sample=64):
    """real-symmetric schedule (csrc/small_sym.cuh): products of commuting symmetric N x N matrices (upper triangle:
    N^2 (N + 1) flops) for cos / sin of Hs, four general real products (2 N^3) for P_seg <- U_n P_seg; real-matrix x
    complex-vector Krylov chains (4 N^2 per order and state), the e_b combination, Im M only, one real trace per
    control.  Orders and classes follow the kernels' tables (sym_tables_upload) and radius bound (sym_radius_bound)."""
    N, L, NT, K, G = p.N, p.L, p.NT, p.K, p.G
    gs = np.unique(np.linspace(0, G - 1, min(G, sample)).astype(int))
    e = np.asarray(eps).reshape(L, NT)
    if p.shape is not None:
        e = e * p.shape
    dt = np.diff(p.tlist)
    H = (p.H0[gs][:, None] + np.einsum("ln,glij->gnij", e, p.Hc[gs])).real
    n1 = np.max(np.sum(np.abs(H), axis=2), axis=2)
    fro = np.sqrt(np.sum(H * H, axis=(2, 3)))
    theta = np.minimum(n1, fro) * dt[None, :]
    econ = econ_on()
    if econ:
        th = econ_theta()
        grad_r = th[:9]
        cls_r = [th[d] for d in SYM_CLS_DEG]
    else:
        grad_r = np.array([0.0, 0.0] + [(2e-17 * _fact(m)) ** (1.0 / m) for m in range(2, 9)])
        cls_r = [min(1.0, (1e-17 * _fact(d + 1)) ** (1.0 / (d + 1))) for d in SYM_CLS_DEG]
    # formation: class, scaling and squaring beyond the top class
    s = np.where(theta > cls_r[-1], np.ceil(np.log2(np.maximum(theta, 1e-300) / cls_r[-1])), 0)
    ths = theta / 2.0 ** s
    cls = np.zeros(theta.shape, dtype=int)
    for c in range(len(cls_r) - 1):
        cls = np.where(ths > cls_r[c], c + 1, cls)
    deg = np.array(SYM_CLS_DEG)[cls]
    ds = (deg - 1) // 2
    horner = np.where(deg % 2 == 0, 1, 0) + 2 * np.maximum(ds - 1, 0)
    symp = float(N * N * (N + 1))
    a_flops = (np.mean((2 + horner + 3 * s) * symp) + 4 * 2.0 * N ** 3 + 2.0 * L * N * N) * G / K
    m = _order_for(theta, grad_r).astype(float)
    nsub = np.where(theta > grad_r[8], np.ceil(theta / grad_r[8]), 1.0)
    c_flops = np.mean(nsub * (12.0 * N * N * m + 10.0 * N * m + 2.0 * N * m * (m - 1))) + 4.0 * L * N * N + N * N
    return dict(formation=float(a_flops), chains=float(8.0 * N * N), contraction=float(c_flops))


def executed_flops_split(p, eps, sample=64, schedule=None):
    """Executed FP64 flops per unit by kernel group: propagator formation (+ segment products), the boundary chains /
    segment fills, the gradient contraction (which carries chi -- and Psi on the Hermitian schedules -- backwards).
    schedule = GrapeEngine.small_schedule() of the measured call (3: real-symmetric kernels)."""
    if schedule == 3:
        return _sym_flops_per_unit(p, eps, sample)
    return _general_flops_per_unit(p, eps, sample)


def executed_flops_per_unit(p, eps, sample=64, schedule=None):
    d = executed_flops_split(p, eps, sample, schedule)
    return d["formation"] + d["chains"] + d["contraction"]


def _general_flops_per_unit(p, eps, sample=64):
    """FP64 flops per (trajectory, step) unit that the small-N / sub-warp kernels execute
    (complex FMA = 8 flops, real-times-complex FMA = 4), following the kernels' own plans:
    propagator formation by Paterson-Stockmeyer Taylor + segment product (per generator-step,
    amortised over the trajectories sharing the generator), the segment fill mat-vecs, and the
    gradient contraction -- Krylov form (small_seg.cuh, m <= 8, no sub-steps, N <= 3, :gradgen)
    or the (1+2L)-mat-vec block recursion with m terms."""
    N, L, NT, K, G = p.N, p.L, p.NT, p.K, p.G
    nrm = _plans(p, eps, sample)
    deg, s = _exp_plan(nrm)
    bs = 4 if N <= 3 else 2
    if bs == 4:
        prods = np.where(deg == 3, 2, 3 + (deg + 1) // 4 - 1)
    else:
        prods = 1 + (deg + 1) // 2 - 1
    seg = p.gb_kind == 0 and N <= 32
    a_flops = (np.mean((prods + s) * 8.0 * N ** 3) + 4.0 * L * N * N + (8.0 * N ** 3 if seg else 0.0)) * G / K
    m, sv = _vec_terms(nrm)
    rec = m * (2.0 ** sv) * (1 + 2 * L) * 8.0 * N * N + L * 8.0 * N
    if seg and N <= 3 and p.gradient_method == 0:
        mm = np.minimum(m, 8)
        eacc = sum(np.where(a < mm, (8 - a) * 4.0 * N, 0.0) for a in range(1, 8)) + 8 * 2.0 * N
        kry = (2 * mm - 1) * 8.0 * N * N + eacc + mm * 8.0 * N * N + mm * 2.0 * N + L * 8.0 * N * N + 4.0 * L * N * N
        fast = (sv == 0) & (m <= 8)
        c_flops = np.mean(np.where(fast, kry, rec))
    else:
        c_flops = np.mean(rec)
    b_flops = 2 * 8.0 * N * N          # forward fill / chain + backward fill / chain mat-vecs
    if seg and N <= 4:
        b_flops = 8.0 * N * N          # chi is carried inside the contraction kernel
    return dict(formation=float(a_flops), chains=float(b_flops), contraction=float(c_flops))


def dense_flops_per_unit(p, eps, form=0):
    """Executed flops per unit of the dense (polynomial-apply) path as (forward, backward sweep, contraction, terms):
    8 N^2 m per state column and Taylor term; + one D Psi product per step each way with g_b.
    form 0, GradGenerator block recursion: the backward sweep has (1 + 2L) operator applications per trajectory
    and term (H^dagger on L+1 blocks, mu_l^dagger on chi) and contains the contraction.
    form 1, Krylov form (csrc/dense_kry.cuh): the backward chain has one column per trajectory, and the
    contraction M_n = sum_{b,k} e_bk ch_bk^dagger costs 8 N^2 m per trajectory-step for all controls together.
    m follows dense_plan() (spectral-norm bound * 1.05; Hermitian generators on the Krylov form: the economised polynomial)."""
    N, L, NT = p.N, p.L, p.NT
    e = np.abs(np.asarray(eps).reshape(L, NT))
    if p.shape is not None:
        e = e * np.abs(p.shape)
    hn = [1.05 * np.linalg.norm(p.H0[0], 2)] + [1.05 * np.linalg.norm(p.Hc[0, l], 2) for l in range(L)]
    nrm = (hn[0] + sum(e[l] * hn[1 + l] for l in range(L))) * np.diff(p.tlist)
    m, sv = _vec_terms(nrm)
    herm = all(np.array_equal(np.asarray(A), np.conj(np.swapaxes(np.asarray(A), -1, -2))) for A in (p.H0, p.Hc))
    if form >= 1 and herm and econ_on() and np.all(nrm <= 1.0):   # Krylov form: economised polynomial (csrc/econ.cuh)
        m = _order_for(nrm, econ_theta())
    terms = float(np.mean(m * 2.0 ** sv))
    gbf = 8.0 * N * N if p.gb_kind else 0.0
    fwd = 8.0 * N * N * terms + gbf
    if form >= 1:
        return fwd, 8.0 * N * N * terms + gbf, 8.0 * N * N * float(np.mean(m)), terms
    return fwd, 8.0 * N * N * terms * (1 + 2 * L) + gbf, 0.0, terms
