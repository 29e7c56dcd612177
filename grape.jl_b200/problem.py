"""Host-side problem container: the hot fields of `GrapeWrk` (reference
src/workspace.jl:78-144) in the flat form the C-ABI descriptor wants
(include/grape_b200.h `grape_b200_problem`)."""
from __future__ import annotations

import numpy as np

SM, RE, SS, HOST = 0, 1, 2, 3
GRADGEN, TAYLOR = 0, 1
JA_NONE, JA_FLUENCE = 0, 1
GB_NONE, GB_QUADFORM = 0, 1
PATH_AUTO, PATH_SMALL, PATH_WARP, PATH_DENSE, PATH_SMALL_CHAIN, PATH_WARP_CHAIN = 0, 1, 2, 3, 4, 5


class GrapeProblem:
    """K trajectories x N levels x L controls x NT intervals.

    H0 [G,N,N], Hc [G,L,N,N] (row/col = matrix indices, i.e. NumPy layout; the
    ctypes layer transposes to the column-major order the ABI specifies),
    psi0/tgt [K,N], gen_of_traj [K].  pulse layout: eps[l*NT + n]
    (reference src/workspace.jl:159-162)."""

    def __init__(self, tlist, H0, Hc, psi0, tgt, gen_of_traj=None, shape=None,
                 weights=None, functional=SM, gradient_method=GRADGEN,
                 ja_kind=JA_NONE, lambda_a=1.0, gb_kind=GB_NONE, lambda_b=1.0,
                 gb_D=None, chi_min_norm=1e-100, taylor_max_order=100,
                 taylor_tolerance=1e-16, taylor_check_convergence=True,
                 K_global=None, path=PATH_AUTO, name=""):
        self.tlist = np.ascontiguousarray(tlist, dtype=np.float64)
        H0 = np.asarray(H0, dtype=np.complex128)
        if H0.ndim == 2:
            H0 = H0[None]
        Hc = np.asarray(Hc, dtype=np.complex128)
        if Hc.ndim == 3:
            Hc = Hc[None]
        if Hc.ndim != 4 or Hc.shape[1] == 0:
            # reference src/workspace.jl:155-157
            raise ValueError("no controls in trajectories: cannot optimize")
        self.H0 = np.ascontiguousarray(H0)
        self.Hc = np.ascontiguousarray(Hc)
        self.psi0 = np.ascontiguousarray(psi0, dtype=np.complex128)
        self.tgt = np.ascontiguousarray(tgt, dtype=np.complex128)
        self.K, self.N = self.psi0.shape
        self.G, self.L = self.Hc.shape[0], self.Hc.shape[1]
        self.NT = len(self.tlist) - 1
        if gen_of_traj is None:
            gen_of_traj = np.arange(self.K) if (self.G == self.K and self.G > 1) \
                else np.zeros(self.K, dtype=np.int32)
        self.gen_of_traj = np.ascontiguousarray(gen_of_traj, dtype=np.int32)
        self.shape = None if shape is None else \
            np.ascontiguousarray(shape, dtype=np.float64).reshape(self.L, self.NT)
        self.weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.functional = int(functional)
        self.gradient_method = int(gradient_method)
        self.ja_kind, self.lambda_a = int(ja_kind), float(lambda_a)
        self.gb_kind, self.lambda_b = int(gb_kind), float(lambda_b)
        if gb_D is not None:
            gb_D = np.asarray(gb_D, dtype=np.complex128)
            if gb_D.ndim == 2:
                gb_D = gb_D[None]
            gb_D = np.ascontiguousarray(gb_D)
        self.gb_D = gb_D
        self.chi_min_norm = float(chi_min_norm)
        self.taylor_max_order = int(taylor_max_order)
        self.taylor_tolerance = float(taylor_tolerance)
        self.taylor_check_convergence = bool(taylor_check_convergence)
        self.K_global = self.K if K_global is None else int(K_global)
        self.path = int(path)
        self.name = name
        assert self.H0.shape == (self.G, self.N, self.N)
        assert self.Hc.shape == (self.G, self.L, self.N, self.N)
        assert self.tgt.shape == (self.K, self.N)

    def shard(self, rank, world):
        """Contiguous block of trajectories for `rank` of `world`;
        generators that no local trajectory uses are dropped."""
        lo = (self.K * rank) // world
        hi = (self.K * (rank + 1)) // world
        ks = np.arange(lo, hi)
        gens, inv = np.unique(self.gen_of_traj[ks], return_inverse=True)
        D = self.gb_D
        if D is not None and D.shape[0] > 1:
            D = D[ks]
        return GrapeProblem(
            self.tlist, self.H0[gens], self.Hc[gens], self.psi0[ks], self.tgt[ks],
            gen_of_traj=inv.astype(np.int32), shape=self.shape,
            weights=None if self.weights is None else self.weights[ks],
            functional=self.functional, gradient_method=self.gradient_method,
            ja_kind=self.ja_kind, lambda_a=self.lambda_a, gb_kind=self.gb_kind,
            lambda_b=self.lambda_b, gb_D=D, chi_min_norm=self.chi_min_norm,
            taylor_max_order=self.taylor_max_order, taylor_tolerance=self.taylor_tolerance,
            taylor_check_convergence=self.taylor_check_convergence,
            K_global=self.K_global, path=self.path, name=f"{self.name}[{rank}/{world}]")
