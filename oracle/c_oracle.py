"""ctypes wrapper of oracle/grape_oracle_c.c -- TEST / BASELINE INFRASTRUCTURE ONLY.

Used by tests (validated against oracle/grape_oracle.py) and by bench.py's
`cpu_baseline` / `--impl reference` legs as the timed multi-threaded CPU
restatement of the reference algorithm.  Never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgrape_oracle.so")
SRC = os.path.join(HERE, "grape_oracle_c.c")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        # -march=native is avoided: the .so is built here and shipped to the GPU box
        cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-std=c11", "-shared", "-o", SO, SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return SO


class _P(C.Structure):
    _fields_ = [("K", C.c_int), ("N", C.c_int), ("L", C.c_int), ("NT", C.c_int), ("G", C.c_int),
                ("K_global", C.c_int), ("functional", C.c_int), ("ja_kind", C.c_int),
                ("gb_kind", C.c_int), ("gb_nD", C.c_int), ("lambda_a", C.c_double),
                ("lambda_b", C.c_double), ("tlist", C.c_void_p), ("gen", C.c_void_p),
                ("H0", C.c_void_p), ("Hc", C.c_void_p), ("shape", C.c_void_p), ("psi0", C.c_void_p),
                ("tgt", C.c_void_p), ("weights", C.c_void_p), ("gb_D", C.c_void_p),
                ("k_count", C.c_int), ("nt_count", C.c_int), ("nthreads", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.grape_oracle_eval_fg.restype = C.c_int
        _lib.grape_oracle_eval_fg.argtypes = [C.POINTER(_P), C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        _lib.grape_oracle_max_threads.restype = C.c_int
        _lib.grape_oracle_expm.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def max_threads():
    return lib().grape_oracle_max_threads()


def expm(A):
    A = np.ascontiguousarray(A, dtype=np.complex128)
    E = np.zeros_like(A)
    lib().grape_oracle_expm(A.shape[0], A.ctypes.data, E.ctypes.data)
    return E


def evaluate_gradient(p, pulsevals, k_count=None, nt_count=None, nthreads=0, want_grad=True):
    """p: any object with the GrapeProblem/OracleProblem attribute names."""
    L_ = lib()
    keep = []

    def ptr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    K, N, L, NT, G = p.K, p.N, p.L, p.NT, p.G
    d = _P()
    d.K, d.N, d.L, d.NT, d.G, d.K_global = K, N, L, NT, G, p.K_global
    d.functional, d.ja_kind, d.gb_kind = p.functional, p.ja_kind, p.gb_kind
    d.lambda_a, d.lambda_b = p.lambda_a, p.lambda_b
    d.tlist = ptr(p.tlist, np.float64)
    d.gen = ptr(p.gen_of_traj, np.int32)
    d.H0, d.Hc = ptr(p.H0, np.complex128), ptr(p.Hc, np.complex128)
    d.shape = ptr(p.shape, np.float64)
    d.psi0, d.tgt = ptr(p.psi0, np.complex128), ptr(p.tgt, np.complex128)
    d.weights = ptr(getattr(p, "weights", None), np.float64)
    gbD = getattr(p, "gb_D", None)
    d.gb_D = ptr(gbD, np.complex128)
    d.gb_nD = 1 if gbD is None else np.asarray(gbD).reshape(-1, N, N).shape[0]
    d.k_count = K if k_count is None else int(k_count)
    d.nt_count = NT if nt_count is None else int(nt_count)
    d.nthreads = int(nthreads)
    eps = np.ascontiguousarray(pulsevals, dtype=np.float64)
    storage = np.zeros((d.k_count, NT + 1, N), dtype=np.complex128)
    Gout = np.zeros(L * NT)
    Jp = np.zeros(3)
    tau = np.zeros(K, dtype=np.complex128)
    rc = L_.grape_oracle_eval_fg(C.byref(d), eps.ctypes.data, storage.ctypes.data,
                                 Gout.ctypes.data if want_grad else None, Jp.ctypes.data, tau.ctypes.data)
    return dict(J=float(Jp.sum()), J_parts=Jp, G=Gout, tau=tau, storage=storage, rc=rc)
