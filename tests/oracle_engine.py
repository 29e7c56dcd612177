"""Oracle-backed stand-in for GrapeEngine (same method surface), so the host
logic (optimize loop, sharding) can be tested on CPU. Test infrastructure only."""
import numpy as np

from oracle import grape_oracle as go


class OracleEngine:
    def __init__(self, problem, device=0):
        self.problem = problem
        self.op = go.from_problem(problem)
        self.K, self.N, self.L, self.NT = problem.K, problem.N, problem.L, problem.NT
        self.J_parts = np.zeros(3)
        self.tau_vals = np.zeros(problem.K, dtype=np.complex128)
        LNT = problem.L * problem.NT
        self.grad_J_Tb, self.grad_J_a = np.zeros(LNT), np.zeros(LNT)
        self.sums = np.zeros(4)
        self._last = None

    def evaluate_functional(self, x):
        r = go.evaluate_functional(self.op, x)
        self._last = r
        self.J_parts[:] = r["J_parts"]
        self.tau_vals[:] = r["tau"]
        return r["J"]

    def evaluate_gradient(self, G, x):
        r = go.evaluate_gradient(self.op, x)
        self._last = r
        G[:] = r["G"]
        self.J_parts[:] = r["J_parts"]
        self.tau_vals[:] = r["tau"]
        self.grad_J_Tb[:] = r["grad_J_Tb"]
        self.grad_J_a[:] = r["grad_J_a"]
        return r["J"]

    def final_states(self):
        return self._last["final_states"]

    # amplitude mode (non-linear controls): pulsevals := amplitudes, control derivatives scaled by dampl
    def evaluate_functional_amplitudes(self, ampl):
        self.op.dampl = np.ones((self.L, self.NT))
        try:
            return self.evaluate_functional(ampl)
        finally:
            self.op.dampl = None

    def evaluate_gradient_amplitudes(self, G_slots, ampl, dampl):
        self.op.dampl = np.asarray(dampl, dtype=np.float64).reshape(self.L, self.NT)
        try:
            return self.evaluate_gradient(G_slots, ampl)
        finally:
            self.op.dampl = None

    # split protocol: emulate with the oracle's sigma_reduce hook
    def forward(self, x):
        self._x = np.array(x, dtype=np.float64)
        r = go.evaluate_functional(self.op, x, sigma_reduce=lambda v: v,
                                   jt_host=(lambda fs: 0.0) if self.op.functional == go.HOST else None)
        self._last = r
        w = self.op.weights
        self.tau_vals[:] = r["tau"]
        s = np.sum(w * r["tau"])
        self.sums[:] = [s.real, s.imag, np.sum(w * np.abs(r["tau"]) ** 2), np.sum(r["J_b_trajectory"])]
        return self.sums.copy()

    def backward_chi(self, chiT, G_partial):
        """host functional: chi_k(T) supplied by the caller (GrapeEngine.backward_chi)"""
        chi = np.asarray(chiT, dtype=np.complex128)
        r = go.evaluate_gradient(self.op, self._x, jt_host=lambda fs: 0.0, chi_host=lambda fs: chi)
        G_partial[:] = r["grad_J_Tb"]
        self.grad_J_a[:] = r["grad_J_a"]
        return float(np.sum(r["J_b_trajectory"]))

    def backward(self, sums_global, G_partial):
        sig = complex(sums_global[0], sums_global[1])

        def red(v):
            if np.ndim(v) == 0:      # scalar ingredient of the functional
                return sig if self.op.functional in (go.SM, go.RE) else complex(sums_global[2], 0.0)
            return v                 # gradient: stays a local partial
        r = go.evaluate_gradient(self.op, self._x, sigma_reduce=red)
        G_partial[:] = r["grad_J_Tb"]
        self.grad_J_a[:] = r["grad_J_a"]
        self.J_parts[:] = r["J_parts"]
        if self.op.gb_kind:
            self.J_parts[2] = self.op.lambda_b * sums_global[3]
        return self.J_parts.copy()


class COracleEngine(OracleEngine):
    """Same surface, backed by the OpenMP C restatement (oracle/grape_oracle_c.c): for host-loop tests whose
    problems are too large for the Python loops (GradGenerator path, built-in functionals)."""

    def _run(self, x, want_grad):
        from oracle import c_oracle as co
        r = co.evaluate_gradient(self.problem, x, want_grad=want_grad)
        assert r["rc"] == 0
        r["final_states"] = r["storage"][:, -1, :].copy()
        self._last = r
        self.J_parts[:] = r["J_parts"]
        self.tau_vals[:] = r["tau"]
        return r

    def evaluate_functional(self, x):
        return self._run(x, False)["J"]

    def evaluate_gradient(self, G, x):
        r = self._run(x, True)
        G[:] = r["G"]
        dt = np.diff(self.problem.tlist)
        ga = (2.0 * np.asarray(x).reshape(self.L, self.NT) * dt[None, :]).reshape(-1) if self.problem.ja_kind else 0.0
        self.grad_J_a[:] = ga
        self.grad_J_Tb[:] = r["G"] - (self.problem.lambda_a * ga if self.problem.ja_kind else 0.0)
        return r["J"]
