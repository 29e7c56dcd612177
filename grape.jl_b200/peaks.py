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


def executed_flops_per_unit(p, eps, sample=64):
    """Real FP64 flops per (trajectory, step) unit executed by the small-N / warp
    kernels (complex FMA = 8 flops): propagator formation by Paterson-Stockmeyer
    Taylor (per generator-step, amortised over the trajectories sharing it),
    two chain mat-vecs, and the (1+2L)-mat-vec block recursion with m terms."""
    N, L, NT, K, G = p.N, p.L, p.NT, p.K, p.G
    gs = np.unique(np.linspace(0, G - 1, min(G, sample)).astype(int))
    e = np.asarray(eps).reshape(L, NT)
    if p.shape is not None:
        e = e * p.shape
    dt = np.diff(p.tlist)
    H = p.H0[gs][:, None] + np.einsum("ln,glij->gnij", e, p.Hc[gs])
    nrm = np.max(np.sum(np.abs(H.real) + np.abs(H.imag), axis=2), axis=2) * dt[None, :]
    deg, s = _exp_plan(nrm)
    bs = 4 if N <= 3 else 2
    if bs == 4:
        prods = np.where(deg == 3, 2, 3 + (deg + 1) // 4 - 1)
    else:
        prods = 1 + (deg + 1) // 2 - 1
    a_flops = np.mean((prods + s) * 8.0 * N ** 3) * G / K
    m, sv = _vec_terms(nrm)
    c_flops = np.mean(m * (2.0 ** sv) * (1 + 2 * L) * 8.0 * N * N) + L * 8.0 * N
    b_flops = 2 * 8.0 * N * N
    return float(a_flops + b_flops + c_flops)
