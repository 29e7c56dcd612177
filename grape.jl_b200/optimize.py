"""Host-side mirror of the reference's `GRAPE.optimize(trajectories, tlist; ...)`
(src/optimize.jl:63-144), `GrapeWrk` set-up (src/workspace.jl:147-362),
`GrapeResult` (src/result.jl:43-110), `update_result!` / `finalize_result!`
(src/optimize.jl:185-228) and the L-BFGS-B reverse-communication loop
(ext/GRAPELBFGSBExt.jl:18-147).

Only the gradient evaluation `fg!(F, G, x)` is replaced: it is one call into
the CUDA engine (GrapeEngine).  The optimizer step stays on the host and is the
same L-BFGS-B 3.0 algorithm the reference links (SciPy's `setulb`), driven with
the reference's exact bound encoding (including its `nbd = 3, u = +Inf` quirk,
ext/GRAPELBFGSBExt.jl:52-63)."""
from __future__ import annotations

import datetime as _dt

import numpy as np

from .configs import discretize_on_midpoints as _disc_mid
from .problem import (GrapeProblem, SM, RE, SS, HOST, GRADGEN, TAYLOR, JA_NONE, JA_FLUENCE,
                      GB_NONE, GB_QUADFORM, PATH_AUTO)


# ---------------------------------------------------------------------------
# problem description objects (QuantumControl.Trajectory, QuantumPropagators.hamiltonian)
# ---------------------------------------------------------------------------
class Control:
    """A control field: a callable eps(t) or an array of values on the intervals of tlist."""

    def __init__(self, func_or_values):
        self.v = func_or_values

    def on_midpoints(self, tlist):
        if callable(self.v):
            return _disc_mid(np.vectorize(self.v, otypes=[float]), tlist)
        v = np.asarray(self.v, dtype=np.float64)
        NT = len(tlist) - 1
        if v.shape == (NT,):
            return v.copy()
        if v.shape == (NT + 1,):   # values on the points of tlist: invert `discretize`
            out = np.zeros(NT)
            out[0] = v[0]
            for i in range(1, NT - 1):
                out[i] = 2 * v[i] - out[i - 1]
            out[-1] = v[-1]
            return out
        raise ValueError("control array must have NT or NT+1 values")


class ShapedAmplitude:
    """a(t) = S(t) * eps(t)  (reference docs/src/tutorial.md:75-107)."""

    def __init__(self, control, shape):
        self.control = control if isinstance(control, Control) else Control(control)
        self.shape = shape


class NonlinearAmplitude:
    """a(t) = f(eps(t), t) with derivative df = d f / d eps -- a control that enters the generator non-linearly
    (reference `get_control_derivs`, src/workspace.jl:283-285: mu = dH/d eps is then itself time dependent and is
    evaluated per step, src/optimize.jl:946-951)."""

    def __init__(self, control, func, dfunc):
        self.control = control if isinstance(control, Control) else Control(control)
        self.func, self.dfunc = func, dfunc


def _control_of(c):
    return c.control if isinstance(c, (ShapedAmplitude, NonlinearAmplitude)) else c


class Generator:
    def __init__(self, H0, terms):
        self.H0 = np.asarray(H0, dtype=np.complex128)
        self.terms = terms   # list of (operator, Control | ShapedAmplitude)


def hamiltonian(H0, *terms):
    """`hamiltonian(H0, (H1, eps1), (H2, eps2), ...)` as in the reference README.md:40-44."""
    out = []
    for op, c in terms:
        if not isinstance(c, (Control, ShapedAmplitude, NonlinearAmplitude)):
            c = Control(c)
        out.append((np.asarray(op, dtype=np.complex128), c))
    return Generator(H0, out)


class Trajectory:
    def __init__(self, initial_state, generator, target_state=None, weight=1.0):
        self.initial_state = np.asarray(initial_state, dtype=np.complex128)
        self.generator = generator
        self.target_state = None if target_state is None else np.asarray(target_state, dtype=np.complex128)
        self.weight = float(weight)


# ---------------------------------------------------------------------------
# functionals (QuantumControl.Functionals)
# ---------------------------------------------------------------------------
class _BuiltinJT:
    def __init__(self, kind, name):
        self.kind, self.__name__ = kind, name

    def __call__(self, states, trajectories, tau=None):
        K = len(trajectories)
        if tau is None:
            tau = np.array([np.vdot(t.target_state, s) for t, s in zip(trajectories, states)])
        w = np.array([t.weight for t in trajectories])
        if self.kind == SM:
            return 1.0 - abs(np.sum(w * tau) / K) ** 2
        if self.kind == RE:
            return 1.0 - np.real(np.sum(w * tau)) / K
        return 1.0 - np.sum(w * np.abs(tau) ** 2) / K


J_T_sm = _BuiltinJT(SM, "J_T_sm")
J_T_re = _BuiltinJT(RE, "J_T_re")
J_T_ss = _BuiltinJT(SS, "J_T_ss")


# ---------------------------------------------------------------------------
# gate functionals: J_T as a function of the achieved gate U_L in the logical subspace
# (QuantumControl.Functionals.gate_functional / make_gate_chi; reference docs/src/background.md:552-610)
# ---------------------------------------------------------------------------
def logical_gate(states, basis):
    """(U_L)_ij = <phi_i | Psi_j(T)>  (docs/src/background.md:556-562); one trajectory per logical basis state."""
    Phi = np.asarray(basis, dtype=np.complex128)          # [d, N]
    Psi = np.asarray(states, dtype=np.complex128)         # [d, N]
    return Phi.conj() @ Psi.T


def gate_functional(J_T_U, basis=None):
    """`J_T(states, trajectories)` from a functional of the gate, `J_T_U(U_L)`.  `basis`: the logical basis states
    (default: the trajectories' initial states, as the reference assumes)."""
    def J_T(states, trajectories, tau=None):
        B = [t.initial_state for t in trajectories] if basis is None else basis
        return float(J_T_U(logical_gate(states, B)))
    J_T.__name__ = getattr(J_T_U, "__name__", "J_T_U") + "_gate"
    return J_T


def make_gate_chi(grad_J_T_U, basis=None):
    """`chi(states, trajectories)` for a gate functional from its gradient `grad_J_T_U(U_L)` = nabla_U J_T
    (= 2 dJ_T/dU*, the reference's complex-gradient convention):
        |chi_k(T)> = -1/2 sum_i (nabla_U J_T)_ik |phi_i>        (docs/src/background.md:600-606)
    The reference obtains nabla_U J_T by automatic differentiation (make_gate_chi); here it is supplied."""
    def chi(states, trajectories, tau=None):
        B = np.asarray([t.initial_state for t in trajectories] if basis is None else basis, dtype=np.complex128)
        g = np.asarray(grad_J_T_U(logical_gate(states, B)), dtype=np.complex128)   # [d, d], index (i, k)
        return -0.5 * (g.T @ B)                                                    # row k = sum_i g[i, k] phi_i
    return chi


def J_a_fluence(pulsevals, tlist):
    dt = np.diff(tlist)
    e = np.asarray(pulsevals).reshape(-1, len(dt))
    return float(np.sum(e * e * dt[None, :]))


class QuadraticForm:
    """g_b(Psi) = <Psi|D|Psi>; xi = -D Psi (reference test/test_state_running_cost.jl:17-65)."""

    def __init__(self, D):
        self.D = np.asarray(D, dtype=np.complex128)

    def __call__(self, psi, *args):
        return float(np.real(np.vdot(psi, self.D @ psi)))


def discretize(vals_on_intervals, tlist):
    """Interval values -> values on the points of tlist (used by finalize_result!, src/optimize.jl:226)."""
    v = np.asarray(vals_on_intervals, dtype=np.float64)
    out = np.zeros(len(tlist))
    out[0], out[-1] = v[0], v[-1]
    out[1:-1] = 0.5 * (v[:-1] + v[1:])
    return out


# ---------------------------------------------------------------------------
# GrapeResult (src/result.jl:43-110)
# ---------------------------------------------------------------------------
class GrapeResult:
    def __init__(self, tlist, guess_pulses, K, N, iter_start=0, iter_stop=5000):
        self.tlist = np.asarray(tlist, dtype=np.float64)
        self.iter_start, self.iter_stop, self.iter = iter_start, iter_stop, iter_start
        self.secs = 0.0
        self.tau_vals = np.zeros(K, dtype=np.complex128)
        self.J_T = self.J_T_prev = 0.0
        self.J_a = self.J_a_prev = 0.0
        self.J_b = self.J_b_prev = 0.0
        self.guess_controls = [discretize(g, tlist) for g in guess_pulses]
        self.optimized_controls = [g.copy() for g in self.guess_controls]
        self.states = np.zeros((K, N), dtype=np.complex128)
        self.start_local_time = self.end_local_time = _dt.datetime.now()
        self.records = []
        self.converged = False
        self.f_calls = self.fg_calls = 0
        self.message = "in progress"

    def __repr__(self):
        return f"GrapeResult<{self.message}>"


# ---------------------------------------------------------------------------
# GrapeWrk host part (src/workspace.jl:147-362)
# ---------------------------------------------------------------------------
def get_controls(trajectories):
    controls = []
    for traj in trajectories:
        for _, c in traj.generator.terms:
            ctrl = _control_of(c)
            if not any(ctrl is x for x in controls):
                controls.append(ctrl)
    return controls


class AmplitudeSlots:
    """Amplitude slots of a problem: one per distinct (control, amplitude) pair that occurs in the generators.

    In the reference the amplitude (shape, non-linearity) belongs to each TERM of each generator.  The engine's
    descriptor has L operator slots per generator; a slot is therefore (control, amplitude kind): 'lin' (a = eps),
    ('shape', S) (a = S(t) eps) or ('nl', amplitude object) (a = f(eps, t)).  If every control has exactly one slot
    and none is non-linear, the slots ARE the controls (`simple`): the engine differentiates with respect to the
    pulse values directly (descriptor `shape`).  Otherwise the engine runs in amplitude mode: the host evaluates
    a and da/d eps per slot and step and adds the slot gradients of each control (chain rule)."""

    def __init__(self, controls, tlist):
        from .configs import midpoints
        self.controls, self.tm, self.NT = controls, midpoints(tlist), len(tlist) - 1
        self.slots = []       # (control index, kind, payload)

    def index_of(self, amp):
        ctrl = _control_of(amp)
        ci = next(i for i, x in enumerate(self.controls) if x is ctrl)
        if isinstance(amp, ShapedAmplitude):
            sv = np.vectorize(amp.shape, otypes=[float])(self.tm) if callable(amp.shape) else \
                np.asarray(amp.shape, dtype=float) * np.ones(self.NT)
            kind, payload = "shape", sv
        elif isinstance(amp, NonlinearAmplitude):
            kind, payload = "nl", amp
        else:
            kind, payload = "lin", None
        for i, (c, k, pl) in enumerate(self.slots):
            if c == ci and k == kind and (k == "lin" or (k == "shape" and np.array_equal(pl, payload)) or
                                          (k == "nl" and pl is payload)):
                return i
        self.slots.append((ci, kind, payload))
        return len(self.slots) - 1

    @property
    def simple(self):
        cs = [c for c, _, _ in self.slots]
        return all(k != "nl" for _, k, _ in self.slots) and sorted(cs) == list(range(len(self.controls)))

    def control_order(self):
        """permutation slot -> position such that slot i of the `simple` case is control i"""
        return [next(i for i, (c, _, _) in enumerate(self.slots) if c == ci) for ci in range(len(self.controls))]

    def shape_array(self):
        """[L, NT] descriptor shape of the `simple` case (None if no control is shaped)"""
        if all(k == "lin" for _, k, _ in self.slots):
            return None
        out = np.ones((len(self.controls), self.NT))
        for c, k, pl in self.slots:
            if k == "shape":
                out[c] = pl
        return out

    def amplitudes(self, pulsevals):
        """(ampl, dampl), each [n_slots * NT]: a_i(eps_c(i),n, t_n) and d a_i / d eps_c(i),n"""
        eps = np.asarray(pulsevals, dtype=np.float64).reshape(len(self.controls), self.NT)
        a = np.zeros((len(self.slots), self.NT))
        da = np.zeros_like(a)
        for i, (c, k, pl) in enumerate(self.slots):
            if k == "lin":
                a[i], da[i] = eps[c], 1.0
            elif k == "shape":
                a[i], da[i] = pl * eps[c], pl
            else:
                a[i] = [pl.func(e, t) for e, t in zip(eps[c], self.tm)]
                da[i] = [pl.dfunc(e, t) for e, t in zip(eps[c], self.tm)]
        return a.reshape(-1), da.reshape(-1)

    def to_controls(self, G_slots):
        """chain rule: dJ/d eps_c = sum of the slot gradients of control c"""
        Gs = np.asarray(G_slots).reshape(len(self.slots), self.NT)
        out = np.zeros((len(self.controls), self.NT))
        for i, (c, _, _) in enumerate(self.slots):
            out[c] += Gs[i]
        return out.reshape(-1)


def build_problem(trajectories, tlist, controls, functional, gradient_method=GRADGEN, ja_kind=JA_NONE,
                  lambda_a=1.0, g_b=None, lambda_b=1.0, **kw):
    """Flatten trajectories into the ABI descriptor's arrays; trajectories that share
    a generator object share one device generator.  Returns (GrapeProblem, AmplitudeSlots)."""
    tlist = np.asarray(tlist, dtype=np.float64)
    K, N, NT = len(trajectories), len(trajectories[0].initial_state), len(tlist) - 1
    gens, gen_of = [], np.zeros(K, dtype=np.int32)
    for k, traj in enumerate(trajectories):
        for gi, g in enumerate(gens):
            if g is traj.generator:
                gen_of[k] = gi
                break
        else:
            gens.append(traj.generator)
            gen_of[k] = len(gens) - 1
    G = len(gens)
    slots = AmplitudeSlots(controls, tlist)
    term_slot = [[slots.index_of(c) for _, c in g.terms] for g in gens]
    if slots.simple:
        order = slots.control_order()
        pos = {s: i for i, s in enumerate(order)}
        L = len(controls)
        shape = slots.shape_array()
    else:
        pos = {i: i for i in range(len(slots.slots))}
        L = len(slots.slots)
        shape = None
        ja_kind = JA_NONE       # J_a acts on the control values: evaluated on the host in amplitude mode
    H0 = np.zeros((G, N, N), dtype=np.complex128)
    Hc = np.zeros((G, L, N, N), dtype=np.complex128)
    for gi, g in enumerate(gens):
        H0[gi] = g.H0
        for (op, _), si in zip(g.terms, term_slot[gi]):
            Hc[gi, pos[si]] += op
    psi0 = np.stack([t.initial_state for t in trajectories])
    tgt = np.stack([t.target_state if t.target_state is not None else np.zeros(N) for t in trajectories])
    w = np.array([t.weight for t in trajectories])
    gb_kind, D = GB_NONE, None
    if g_b is not None and lambda_b != 0.0:
        if not isinstance(g_b, QuadraticForm):
            raise NotImplementedError(
                "only the built-in quadratic form g_b runs on the device (arbitrary g_b/xi closures "
                "would need every stored state on the host; DESIGN.md, out of scope)")
        gb_kind, D = GB_QUADFORM, g_b.D
    prob = GrapeProblem(tlist, H0, Hc, psi0, tgt, gen_of_traj=gen_of, shape=shape,
                        weights=None if np.all(w == 1.0) else w, functional=functional,
                        gradient_method=gradient_method, ja_kind=ja_kind, lambda_a=lambda_a,
                        gb_kind=gb_kind, lambda_b=lambda_b, gb_D=D, **kw)
    return prob, slots


class GrapeWrk:
    """Host mirror of the reference workspace; device buffers live in `engine`."""

    def __init__(self, trajectories, tlist, kwargs, engine_factory=None):
        trajectories = list(trajectories)
        self.trajectories, self.kwargs = trajectories, dict(kwargs)
        self.tlist = np.asarray(tlist, dtype=np.float64)
        self.controls = get_controls(trajectories)
        if len(self.controls) == 0:
            raise RuntimeError("no controls in trajectories: cannot optimize")   # workspace.jl:155-157
        NT = len(self.tlist) - 1
        kw = self.kwargs
        if "J_T" not in kw:
            raise TypeError("`optimize` for `method=GRAPE` must be passed the functional `J_T`.")  # workspace.jl:298-303
        guess = [c.on_midpoints(self.tlist) for c in self.controls]
        self.pulsevals = np.concatenate(guess)                                   # workspace.jl:159-162
        K, N = len(trajectories), len(trajectories[0].initial_state)
        if "continue_from" in kw:                                                # workspace.jl:167-186
            res = kw["continue_from"]
            res.iter_stop = kw.get("iter_stop", 5000)
            res.converged, res.message = False, "in progress"
            res.start_local_time = _dt.datetime.now()
            self.pulsevals = np.concatenate(
                [Control(c).on_midpoints(res.tlist) for c in res.optimized_controls])
            self.result = res
        else:
            self.result = GrapeResult(self.tlist, guess, K, N, kw.get("iter_start", 0), kw.get("iter_stop", 5000))
        n = len(self.pulsevals)
        self.pulsevals_guess = self.pulsevals.copy()
        self.gradient = np.zeros(n)
        self.grad_J_Tb, self.grad_J_a = np.zeros(n), np.zeros(n)
        self.J_parts = np.zeros(3)
        self.fg_count = np.zeros(2, dtype=np.int64)
        self.upper_bounds = np.full(n, kw.get("upper_bound", np.inf), dtype=np.float64)   # workspace.jl:202-214
        self.lower_bounds = np.full(n, kw.get("lower_bound", -np.inf), dtype=np.float64)
        L = len(self.controls)
        for l, c in enumerate(self.controls):
            opts = kw.get("pulse_options", {}).get(c, {})
            if "upper_bounds" in opts:
                self.upper_bounds[l::L] = opts["upper_bounds"]     # interleaved stride, as the reference
            if "lower_bounds" in opts:
                self.lower_bounds[l::L] = opts["lower_bounds"]
        J_T = kw["J_T"]
        self.J_T_func, self.chi_func = J_T, kw.get("chi")
        functional = J_T.kind if isinstance(J_T, _BuiltinJT) and self.chi_func is None else HOST
        if functional == HOST and self.chi_func is None:
            raise NotImplementedError("a custom J_T needs an explicit `chi` (no automatic differentiation here)")
        self.J_a_func, self.grad_J_a_func = kw.get("J_a"), kw.get("grad_J_a")
        self.lambda_a = kw.get("lambda_a", 1.0)
        ja_kind = JA_NONE
        if self.J_a_func is J_a_fluence and self.grad_J_a_func is None:
            ja_kind = JA_FLUENCE          # built-in device reduction
        elif self.J_a_func is not None and self.grad_J_a_func is None:
            raise NotImplementedError("a custom J_a needs an explicit grad_J_a")
        self.host_J_a = self.J_a_func is not None and ja_kind == JA_NONE
        gm = kw.get("gradient_method", "gradgen")
        if gm not in ("gradgen", "taylor"):
            raise ValueError(f"Invalid gradient_method={gm!r} ∉ (:gradgen, :taylor)")
        self.problem, self.slots = build_problem(
            trajectories, self.tlist, self.controls, functional,
            gradient_method=GRADGEN if gm == "gradgen" else TAYLOR, ja_kind=ja_kind,
            lambda_a=self.lambda_a, g_b=kw.get("g_b"), lambda_b=kw.get("lambda_b", 1.0),
            chi_min_norm=kw.get("chi_min_norm", 1e-100),
            taylor_max_order=kw.get("taylor_grad_max_order", 100),
            taylor_tolerance=kw.get("taylor_grad_tolerance", 1e-16),
            taylor_check_convergence=kw.get("taylor_grad_check_convergence", True),
            path=kw.get("path", PATH_AUTO))
        self.amplitude_mode = not self.slots.simple
        if self.amplitude_mode:
            if functional == HOST:
                raise NotImplementedError("a custom chi together with non-linear / per-term amplitudes is not supported")
            if ja_kind == JA_FLUENCE:             # the device never sees the control values in amplitude mode
                self.host_J_a = True
                self.grad_J_a_func = lambda x, tl: (2.0 * np.asarray(x).reshape(-1, len(tl) - 1)
                                                    * np.diff(tl)[None, :]).reshape(-1)
        if engine_factory is None:
            from .engine import GrapeEngine
            engine_factory = lambda prob: GrapeEngine(prob, device=kw.get("device", 0))
        self.engine = engine_factory(self.problem)

    # -- the two closures of src/optimize.jl:98-111 -------------------------------
    def evaluate_functional(self, pulsevals, count_call=True):
        e = self.engine
        if self.problem.functional == HOST:
            e.forward(pulsevals)
            states = e.final_states()
            self.J_parts[:] = 0.0
            self.J_parts[0] = self.J_T_func(list(states), self.trajectories)
            self.J_parts[2] = self.problem.lambda_b * e.sums[3] if self.problem.gb_kind else 0.0
        elif self.amplitude_mode:
            e.evaluate_functional_amplitudes(self.slots.amplitudes(pulsevals)[0])
            self.J_parts[:] = e.J_parts
        else:
            e.evaluate_functional(pulsevals)
            self.J_parts[:] = e.J_parts
        if self.host_J_a:
            self.J_parts[1] = self.lambda_a * self.J_a_func(pulsevals, self.tlist)
        self.result.tau_vals[:] = e.tau_vals
        if count_call:
            self.result.f_calls += 1
            self.fg_count[1] += 1
        return float(np.sum(self.J_parts))

    def evaluate_gradient(self, G, pulsevals):
        e = self.engine
        self.result.fg_calls += 1
        self.fg_count[0] += 1
        if self.problem.functional == HOST:
            e.forward(pulsevals)
            states = e.final_states()
            self.J_parts[:] = 0.0
            self.J_parts[0] = self.J_T_func(list(states), self.trajectories)
            chi = np.asarray(self.chi_func(list(states), self.trajectories), dtype=np.complex128)
            jb = e.backward_chi(chi, self.grad_J_Tb)
            self.J_parts[2] = self.problem.lambda_b * jb if self.problem.gb_kind else 0.0
            G[:] = self.grad_J_Tb
            self.grad_J_a[:] = e.grad_J_a
            if self.problem.ja_kind:
                self.J_parts[1] = self.lambda_a * J_a_fluence(pulsevals, self.tlist)
                G += self.lambda_a * self.grad_J_a
        elif self.amplitude_mode:
            # non-linear / per-term amplitudes: the host evaluates a and da/d eps per slot (get_control_derivs,
            # src/workspace.jl:283-285; evaluate(mu), src/optimize.jl:946-951) and applies the chain rule
            ampl, dampl = self.slots.amplitudes(pulsevals)
            G_slots = np.zeros_like(ampl)
            e.evaluate_gradient_amplitudes(G_slots, ampl, dampl)
            self.J_parts[:] = e.J_parts
            self.grad_J_Tb[:] = self.slots.to_controls(G_slots)
            G[:] = self.grad_J_Tb
            self.grad_J_a[:] = 0.0
        else:
            e.evaluate_gradient(G, pulsevals)
            self.J_parts[:] = e.J_parts
            self.grad_J_Tb[:] = e.grad_J_Tb
            self.grad_J_a[:] = e.grad_J_a
        if self.host_J_a:                                        # src/optimize.jl:1004-1011
            self.J_parts[1] = self.lambda_a * self.J_a_func(pulsevals, self.tlist)
            self.grad_J_a[:] = self.grad_J_a_func(pulsevals, self.tlist)
            G += self.lambda_a * self.grad_J_a
        self.result.tau_vals[:] = e.tau_vals
        return float(np.sum(self.J_parts))


def update_result(wrk, i):
    """src/optimize.jl:185-216"""
    res = wrk.result
    res.states[:] = wrk.engine.final_states()
    res.J_T_prev, res.J_T = res.J_T, wrk.J_parts[0]
    res.J_a_prev, res.J_a = res.J_a, wrk.J_parts[1]
    if res.J_a > 0.0:
        res.J_a /= wrk.lambda_a
    res.J_b_prev = res.J_b
    lam_b = wrk.kwargs.get("lambda_b", 1.0)
    res.J_b = wrk.J_parts[2] / lam_b if not (lam_b == 0 and wrk.kwargs.get("g_b") is None) else 0.0
    if i > 0:
        res.iter = i
    if i >= res.iter_stop:
        res.converged = True
        res.message = "Reached maximum number of iterations"
    prev = res.end_local_time
    res.end_local_time = _dt.datetime.now()
    res.secs = (res.end_local_time - prev).total_seconds()


def apply_convergence_check(result, check_convergence):
    """src/optimize.jl:154-182"""
    if result.converged:
        return
    c = check_convergence(result)
    if isinstance(c, bool):
        result.converged = c
        if c:
            result.message = "Convergence check returned true"
    elif isinstance(c, str):
        if c:
            result.converged = True
            result.message = c
    # None or the (mutated) result object: nothing to do


# ---------------------------------------------------------------------------
# L-BFGS-B reverse-communication loop (ext/GRAPELBFGSBExt.jl:18-147)
# ---------------------------------------------------------------------------
def run_lbfgsb(x, fg, lower_bounds, upper_bounds, on_start, on_new_x, m=10, factr=1e1, pgtol=1e-15,
               maxls=20):
    """Drive SciPy's `setulb` (C port of L-BFGS-B 3.0, the algorithm LBFGSB.jl wraps).

    `fg(x, g) -> f` fills `g` in place.  `on_start(g)` is called after the first
    evaluation (FG_START), `on_new_x(g) -> stop?` after each iteration.  `x` is
    updated in place (the reference aliases `wrk.pulsevals`, ext:31).
    Returns the final task message."""
    from scipy.optimize import _lbfgsb
    n = len(x)
    nbd = np.zeros(n, dtype=np.int32)
    lo, up = np.zeros(n), np.zeros(n)
    for i in range(n):                       # ext/GRAPELBFGSBExt.jl:47-64, verbatim logic
        if lower_bounds[i] > -np.inf:
            nbd[i] += 1
            lo[i] = lower_bounds[i]
        if upper_bounds[i] > -np.inf:        # (sic) always true for +Inf: nbd = 3 with u = +Inf
            nbd[i] = 2 if nbd[i] == 1 else 3
            up[i] = upper_bounds[i]
    f = np.array(0.0)
    g = np.zeros(n)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
    iwa = np.zeros(3 * n, dtype=np.int32)
    task, ln_task = np.zeros(2, dtype=np.int32), np.zeros(2, dtype=np.int32)
    lsave, isave, dsave = np.zeros(4, dtype=np.int32), np.zeros(44, dtype=np.int32), np.zeros(29)
    from scipy.optimize._lbfgsb_py import status_messages, task_messages
    while True:
        _lbfgsb.setulb(m, x, lo, up, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave,
                       maxls, ln_task)
        if task[0] == 3:                     # FG_START / FG_LNSRCH   (ext:97-109)
            f = np.array(fg(x, g))
            if task[1] == 301:
                on_start(g)
        elif task[0] == 1:                   # NEW_X  (ext:110-127)
            if on_new_x(g):
                task[0], task[1] = 5, 505    # "STOP: NEW_X -> CONVERGED"
        else:
            return (status_messages.get(int(task[0]), "?") + ": " + task_messages.get(int(task[1]), "")).strip()


def optimize(trajectories, tlist, *, engine_factory=None, **kwargs):
    """`GRAPE.optimize(trajectories, tlist; kwargs...)` (src/optimize.jl:73-144)."""
    if "update_hook" in kwargs or "info_hook" in kwargs:
        raise TypeError("The `update_hook` and `info_hook` arguments have been superseded by the `callback` argument")
    callback = kwargs.get("callback", lambda *a: None)
    check_convergence = kwargs.get("check_convergence", lambda res: res)
    wrk = GrapeWrk(trajectories, tlist, kwargs, engine_factory=engine_factory)
    res = wrk.result

    def fg(x, g):
        return wrk.evaluate_gradient(g, x)

    def record(info):
        if info is not None and len(info) > 0:
            res.records.append(tuple(info))

    def on_start(g):
        wrk.gradient[:] = g
        update_result(wrk, 0)
        record(callback(wrk, 0))
        wrk.fg_count[:] = 0

    def on_new_x(g):
        update_result(wrk, res.iter + 1)
        record(callback(wrk, res.iter))
        wrk.fg_count[:] = 0
        apply_convergence_check(res, check_convergence)
        if not res.converged:
            wrk.pulsevals_guess[:] = wrk.pulsevals
            wrk.gradient[:] = g
        return res.converged

    try:
        msg = run_lbfgsb(wrk.pulsevals, fg, wrk.lower_bounds, wrk.upper_bounds, on_start, on_new_x,
                         m=kwargs.get("lbfgsb_m", 10), factr=kwargs.get("lbfgsb_factr", 1e1),
                         pgtol=kwargs.get("lbfgsb_pgtol", 1e-15))
        if res.message == "in progress":
            res.message = msg
    except Exception as exc:                                    # src/optimize.jl:125-135
        if kwargs.get("rethrow_exceptions", False):
            raise
        res.message = f"Exception: {exc}"
    # finalize_result!  (src/optimize.jl:219-228)
    res.end_local_time = _dt.datetime.now()
    NT = len(res.tlist) - 1
    for l in range(len(wrk.controls)):
        res.optimized_controls[l] = discretize(wrk.pulsevals[l * NT:(l + 1) * NT], res.tlist)
    res.pulsevals = wrk.pulsevals.copy()
    res.wrk = wrk
    return res
