"""GrapeEngine: the device-resident `GrapeWrk` (reference src/workspace.jl:78-362)
plus `evaluate_functional` / `evaluate_gradient!` (reference
src/optimize.jl:696-768, 824-1014) as thin calls through the C-ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .problem import GrapeProblem, HOST


class GrapeError(RuntimeError):
    """Raised with the library's message; mirrors the Julia `error(...)` calls of
    the reference hot path (src/optimize.jl:644-648, 1021-1025; src/workspace.jl:155-157)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_D1 = C.c_double * 1


def _dp(a):
    """double* of a NumPy array for the C-ABI.  `from_buffer` (0.7 us) instead of `ctypes.data_as` (3.7 us): a call
    converts the pulse and gradient arrays every time; read-only or empty arrays take the slow way."""
    if a is None:
        return None
    try:
        return _D1.from_buffer(a)
    except (TypeError, ValueError, BufferError):
        return a.ctypes.data_as(C.POINTER(C.c_double))


def _colmajor(mats):
    """[..., N, N] NumPy matrices -> contiguous array whose last two axes are
    stored column-major (the ABI's Julia layout)."""
    return np.ascontiguousarray(np.swapaxes(mats, -1, -2))


def _descriptor(p: GrapeProblem, device: int, keep: dict):
    """Fill a `grape_b200_problem` from a GrapeProblem; `keep` holds the arrays the pointers refer to."""
    keep["tlist"] = p.tlist
    keep["H0"] = _colmajor(p.H0)
    keep["Hc"] = _colmajor(p.Hc)
    keep["psi0"], keep["tgt"] = p.psi0, p.tgt
    keep["gen"] = p.gen_of_traj
    d = _lib.ProblemDesc()
    d.abi_version = _lib.ABI_VERSION
    d.K, d.N, d.L, d.NT, d.G = p.K, p.N, p.L, p.NT, p.G
    d.K_global, d.device = p.K_global, device
    d.tlist = _dp(keep["tlist"])
    d.gen_of_traj = keep["gen"].ctypes.data_as(C.POINTER(C.c_int32))
    d.H0, d.Hc = _dp(keep["H0"].view(np.float64)), _dp(keep["Hc"].view(np.float64))
    d.shape = _dp(p.shape)
    d.psi0, d.tgt = _dp(p.psi0.view(np.float64)), _dp(p.tgt.view(np.float64))
    d.weights = _dp(p.weights)
    d.functional, d.gradient_method = p.functional, p.gradient_method
    d.ja_kind, d.gb_kind = p.ja_kind, p.gb_kind
    d.lambda_a, d.lambda_b = p.lambda_a, p.lambda_b
    if p.gb_D is not None and p.gb_kind:
        keep["D"] = _colmajor(p.gb_D)
        d.gb_D = _dp(keep["D"].view(np.float64))
        d.gb_nD = p.gb_D.shape[0]
    d.taylor_max_order = p.taylor_max_order
    d.taylor_tolerance = p.taylor_tolerance
    d.taylor_check_convergence = int(p.taylor_check_convergence)
    d.path = p.path
    d.chi_min_norm = p.chi_min_norm
    return d


class GrapeEngine:
    def __init__(self, problem: GrapeProblem, device: int = 0):
        self.lib = _lib.load()
        p = self.problem = problem
        self.K, self.N, self.L, self.NT = p.K, p.N, p.L, p.NT
        self._keep = keep = {}
        d = _descriptor(p, device, keep)
        h = C.c_void_p()
        rc = self.lib.grape_b200_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise GrapeError(rc, (self.lib.grape_b200_last_error(None) or b"").decode())
        self._h = h
        LNT = p.L * p.NT
        # host mirrors of the GrapeWrk fields callbacks read (src/optimize.jl:402-478)
        self.J_parts = np.zeros(3)
        self.tau_vals = np.zeros(p.K, dtype=np.complex128)
        self.grad_J_Tb = np.zeros(LNT)
        self.grad_J_a = np.zeros(LNT)
        # ctypes pointers of the engine-owned result arrays, made once (each conversion costs ~1.5 us of the ~35 us
        # host overhead of a call; the arrays are never re-allocated)
        self._tau_f64 = self.tau_vals.view(np.float64)
        self._pp = (_dp(self.J_parts), _dp(self._tau_f64), _dp(self.grad_J_Tb), _dp(self.grad_J_a))
        self.sums = np.zeros(4)
        self.fg_calls = 0
        self.f_calls = 0

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.grape_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GrapeError(rc, (self.lib.grape_b200_last_error(self._h) or b"").decode())

    @staticmethod
    def _pulses(pulsevals, LNT):
        x = np.ascontiguousarray(pulsevals, dtype=np.float64)
        if x.shape != (LNT,):
            raise ValueError(f"pulsevals must have length L*NT = {LNT}")
        return x

    # -- evaluate_functional (src/optimize.jl:696-768) --------------------------
    def evaluate_functional(self, pulsevals):
        x = self._pulses(pulsevals, self.L * self.NT)
        self._check(self.lib.grape_b200_eval_f(self._h, _dp(x), self._pp[0], self._pp[1]))
        self.f_calls += 1
        return float(self.J_parts[0] + self.J_parts[1] + self.J_parts[2])

    # -- evaluate_gradient! (src/optimize.jl:824-1014) ---------------------------
    def evaluate_gradient(self, G, pulsevals):
        x = self._pulses(pulsevals, self.L * self.NT)
        if not (isinstance(G, np.ndarray) and G.dtype == np.float64 and G.flags.c_contiguous
                and G.shape == x.shape):
            raise ValueError("G must be a contiguous float64 array of length L*NT")
        self._check(self.lib.grape_b200_eval_fg(self._h, _dp(x), _dp(G), *self._pp))
        self.fg_calls += 1
        return float(self.J_parts[0] + self.J_parts[1] + self.J_parts[2])

    # -- amplitude mode: non-linear controls / per-term amplitudes (include/grape_b200.h) ------------------
    def evaluate_functional_amplitudes(self, ampl):
        a = self._pulses(ampl, self.L * self.NT)
        self._check(self.lib.grape_b200_eval_f_amplitudes(self._h, _dp(a), _dp(self.J_parts),
                                                          _dp(self.tau_vals.view(np.float64))))
        self.f_calls += 1
        return float(np.sum(self.J_parts))

    def evaluate_gradient_amplitudes(self, G_slots, ampl, dampl):
        a = self._pulses(ampl, self.L * self.NT)
        da = self._pulses(dampl, self.L * self.NT)
        assert isinstance(G_slots, np.ndarray) and G_slots.dtype == np.float64 and G_slots.shape == a.shape
        self._check(self.lib.grape_b200_eval_fg_amplitudes(self._h, _dp(a), _dp(da), _dp(G_slots), _dp(self.J_parts),
                                                           _dp(self.tau_vals.view(np.float64))))
        self.fg_calls += 1
        return float(np.sum(self.J_parts))

    # -- split form --------------------------------------------------------------
    def forward(self, pulsevals):
        x = self._pulses(pulsevals, self.L * self.NT)
        self._check(self.lib.grape_b200_forward(self._h, _dp(x), _dp(self.tau_vals.view(np.float64)),
                                                _dp(self.sums)))
        return self.sums.copy()

    def backward(self, sums_global, G_partial):
        s = np.ascontiguousarray(sums_global, dtype=np.float64)
        self._check(self.lib.grape_b200_backward(self._h, _dp(s), _dp(G_partial), _dp(self.J_parts),
                                                 _dp(self.grad_J_a)))
        return self.J_parts.copy()

    def backward_chi(self, chiT, G_partial):
        c = np.ascontiguousarray(chiT, dtype=np.complex128)
        assert c.shape == (self.K, self.N)
        jb = C.c_double(0.0)
        self._check(self.lib.grape_b200_backward_chi(self._h, _dp(c.view(np.float64)), _dp(G_partial),
                                                     C.byref(jb), _dp(self.grad_J_a)))
        return jb.value

    # -- workspace read-backs ------------------------------------------------------
    def final_states(self):
        out = np.zeros((self.K, self.N), dtype=np.complex128)
        self._check(self.lib.grape_b200_get_final_states(self._h, _dp(out.view(np.float64))))
        return out

    def stored_states(self, k):
        """fw_storage[k] as an [N, NT+1] array (column n = Psi_k(t_n))."""
        out = np.zeros((self.NT + 1, self.N), dtype=np.complex128)
        self._check(self.lib.grape_b200_get_stored_states(self._h, int(k), _dp(out.view(np.float64))))
        return out.T

    def chi_states(self):
        chi = np.zeros((self.K, self.N), dtype=np.complex128)
        rho = np.zeros(self.K)
        self._check(self.lib.grape_b200_get_chi_states(self._h, _dp(chi.view(np.float64)), _dp(rho)))
        return chi, rho

    def tau_grads(self, k):
        """tau_grads[k] as an [NT, L] complex array (src/workspace.jl:236-237)."""
        out = np.zeros((self.L, self.NT), dtype=np.complex128)
        self._check(self.lib.grape_b200_get_tau_grads(self._h, int(k), _dp(out.view(np.float64))))
        return out.T

    # -- instrumentation -------------------------------------------------------------
    def set_profiling(self, on=True):
        self._check(self.lib.grape_b200_set_profiling(self._h, int(bool(on))))

    def timings(self):
        t = np.zeros(8)
        self._check(self.lib.grape_b200_get_timings(self._h, _dp(t)))
        return dict(formU_ms=t[0], forward_ms=t[1], tau_ms=t[2], backward_ms=t[3], gradient_ms=t[4],
                    d2h_ms=t[5], total_ms=t[6], launches=int(t[7]))

    def launch_count(self):
        return int(self.lib.grape_b200_launch_count(self._h))

    def gradient_form(self):
        """0: GradGenerator block recursion, 1: Krylov form (dense path), 2: Krylov form with two Taylor terms per
        grid barrier in the strip chains served the last gradient call."""
        rc = int(self.lib.grape_b200_gradient_form(self._h))
        if rc < 0:
            self._check(-rc)
        return rc

    def dense_concurrent(self):
        """1: the last gradient call ran the forward sweep and the chi chain of the dense path concurrently."""
        rc = int(self.lib.grape_b200_dense_concurrent(self._h))
        if rc < 0:
            self._check(-rc)
        return rc

    def dense_orders(self):
        """(econ, orders): polynomial degree of every time step in the last call of the dense path's Krylov-form
        schedule; econ = True if the chains summed the economised (Chebyshev-cut) polynomial instead of the Taylor series."""
        out = np.zeros(self.problem.NT, dtype=np.int32)
        rc = int(self.lib.grape_b200_dense_orders(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        if rc < 0:
            self._check(-rc)
        return bool(rc), out

    def small_schedule(self):
        """0: not the segmented small path, 1: general, 2: Hermitian, 3: real-symmetric schedule served the last gradient."""
        rc = int(self.lib.grape_b200_small_schedule(self._h))
        if rc < 0:
            self._check(-rc)
        return rc

    def eval_fg_device(self, d_pulsevals_ptr, d_G_ptr=None, d_J_ptr=None):
        self._check(self.lib.grape_b200_eval_fg_device(self._h, d_pulsevals_ptr, d_G_ptr, d_J_ptr))

    def enqueue_forward(self, d_pulsevals_ptr):
        self._check(self.lib.grape_b200_enqueue_forward(self._h, d_pulsevals_ptr))

    def enqueue_backward(self):
        self._check(self.lib.grape_b200_enqueue_backward(self._h))

    def enqueue_combine(self):
        self._check(self.lib.grape_b200_enqueue_combine(self._h))

    def finish(self):
        self._check(self.lib.grape_b200_finish(self._h))

    def device_ptr(self, which):
        return self.lib.grape_b200_device_ptr(self._h, which)

    def stream(self):
        return self.lib.grape_b200_stream(self._h)

    # -- peer exchange between shards, one process per GPU (include/grape_b200.h, csrc/xchg.cuh) -------
    def xchg_init(self, rank, world):
        """Allocate this shard's exchange buffer; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        self._check(self.lib.grape_b200_xchg_init(self._h, int(rank), int(world), buf))
        return buf.raw

    def xchg_attach(self, handles):
        """`handles`: the IPC handles of ALL ranks in rank order. Afterwards evaluate_* / enqueue_* are collective
        calls that return the global J and gradient on every rank."""
        blob = b"".join(handles)
        assert len(blob) % _lib.IPC_HANDLE_BYTES == 0
        self._check(self.lib.grape_b200_xchg_attach(self._h, C.c_char_p(blob)))

    def xchg_detach(self):
        self._check(self.lib.grape_b200_xchg_detach(self._h))


class MultiGrapeEngine:
    """The WHOLE problem on several GPUs driven by this one process (`grape_b200_multi_*`): same
    `evaluate_functional` / `evaluate_gradient` surface as GrapeEngine."""

    def __init__(self, problem: GrapeProblem, devices):
        self.lib = _lib.load()
        p = self.problem = problem
        self.K, self.N, self.L, self.NT = p.K, p.N, p.L, p.NT
        self._keep = {}
        d = _descriptor(p, 0, self._keep)
        devs = (C.c_int32 * len(devices))(*[int(x) for x in devices])
        h = C.c_void_p()
        rc = self.lib.grape_b200_multi_create(C.byref(d), devs, len(devices), C.byref(h))
        if rc != 0:
            raise GrapeError(rc, (self.lib.grape_b200_multi_last_error(None) or b"").decode())
        self._h = h
        LNT = p.L * p.NT
        self.J_parts = np.zeros(3)
        self.tau_vals = np.zeros(p.K, dtype=np.complex128)
        self.grad_J_Tb, self.grad_J_a = np.zeros(LNT), np.zeros(LNT)
        self.fg_calls = self.f_calls = 0

    def close(self):
        if getattr(self, "_h", None):
            self.lib.grape_b200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GrapeError(rc, (self.lib.grape_b200_multi_last_error(self._h) or b"").decode())

    def size(self):
        return int(self.lib.grape_b200_multi_size(self._h))

    def evaluate_functional(self, pulsevals):
        x = GrapeEngine._pulses(pulsevals, self.L * self.NT)
        self._check(self.lib.grape_b200_multi_eval_f(self._h, _dp(x), _dp(self.J_parts),
                                                     _dp(self.tau_vals.view(np.float64))))
        self.f_calls += 1
        return float(np.sum(self.J_parts))

    def evaluate_gradient(self, G, pulsevals):
        x = GrapeEngine._pulses(pulsevals, self.L * self.NT)
        if not (isinstance(G, np.ndarray) and G.dtype == np.float64 and G.flags.c_contiguous and G.shape == x.shape):
            raise ValueError("G must be a contiguous float64 array of length L*NT")
        self._check(self.lib.grape_b200_multi_eval_fg(
            self._h, _dp(x), _dp(G), _dp(self.J_parts), _dp(self.tau_vals.view(np.float64)),
            _dp(self.grad_J_Tb), _dp(self.grad_J_a)))
        self.fg_calls += 1
        return float(np.sum(self.J_parts))

    def final_states(self):
        out = np.zeros((self.K, self.N), dtype=np.complex128)
        self._check(self.lib.grape_b200_multi_get_final_states(self._h, _dp(out.view(np.float64))))
        return out
