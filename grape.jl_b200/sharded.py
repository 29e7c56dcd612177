"""Trajectory sharding over ranks (one process per GPU).

The trajectory ensemble is split into contiguous blocks, one per rank
(`GrapeProblem.shard`).  Per gradient evaluation there are exactly two tiny
exchanges, the only points where trajectories couple (reference
src/optimize.jl:755-760, 845-855 through the functional, and :574-584 through
the sum over k):
  1. all-reduce(sum) of the 4 partial sums (sum w tau, sum w|tau|^2, sum J_b)
     after the forward sweep;
  2. all-reduce(sum) of the partial gradient [L*NT] after the backward sweep.
For J_T_re / J_T_ss chi_k does not depend on the other trajectories, so (1) is only
needed for the value of J and travels with (2).

Default exchange ("p2p"): the library's own reduction kernels push the partials into
every peer's exchange buffer over NVLink and add them in rank order (csrc/xchg.cuh) --
`torch.distributed` is only used ONCE, at set-up, to pass the 64-byte CUDA IPC handles
around; there is no collective-library launch and no Python between the kernels of a
step.  exchange="nccl" keeps the in-place NCCL all-reduces of round 1 (comparison /
boxes without peer access).  The host-side optimizer step is unchanged and identical
on every rank."""
from __future__ import annotations

import numpy as np


class ShardedGrape:
    """Same `evaluate_gradient(G, x)` / `evaluate_functional(x)` surface as GrapeEngine,
    for a problem sharded over `torch.distributed` ranks.

    `engine_factory(local_problem) -> engine` builds the per-rank engine (the CUDA
    GrapeEngine in production).  Collectives run on `device` tensors when given
    (NCCL) or on CPU tensors (gloo)."""

    def __init__(self, problem, engine_factory, rank=None, world=None, device=None, group=None, exchange="p2p"):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.problem = problem
        self.local = problem.shard(self.rank, self.world)
        self.engine = engine_factory(self.local)
        self.device = device
        LNT = problem.L * problem.NT
        kw = dict(dtype=torch.float64, device=device if device is not None else "cpu")
        self._sums = torch.zeros(4, **kw)
        self._grad = torch.zeros(LNT, **kw)
        self._G_partial = np.zeros(LNT)
        # device-resident pipeline (CUDA engine): pinned staging, one stream, one synchronisation per call
        self._pipe = None
        if device is not None and getattr(device, "type", "cpu") == "cuda" and hasattr(self.engine, "enqueue_forward"):
            dd = dist if self.world > 1 else None
            self._pipe = DevicePipeline(self.engine, dd, group, exchange=exchange)
            self._stream = torch.cuda.ExternalStream(self.engine.stream(), device=device)
            self._h_eps = torch.empty(LNT, dtype=torch.float64).pin_memory()
            self._d_eps = torch.empty(LNT, dtype=torch.float64, device=device)
            self._h_out = torch.empty(LNT + 3, dtype=torch.float64).pin_memory()
            self._J_t = torch.as_tensor(_DevArray(self.engine.device_ptr(4), 3), device=device)
        self.J_parts = np.zeros(3)
        self.grad_J_Tb = np.zeros(LNT)
        self.grad_J_a = np.zeros(LNT)
        self.K_local = self.local.K

    def close(self):
        """Releases the pinned staging buffers while the engine's stream is still alive: torch's
        pinned-host allocator records an event on every stream a block was used on when the block
        is freed, so they must go before `grape_b200_destroy` destroys that stream."""
        if self._pipe is not None:
            self.engine.finish()
            self._pipe.close()
            self._h_eps = self._h_out = self._d_eps = self._J_t = None
            self._pipe = None
            self._stream = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _allreduce(self, buf, host):
        t = self._torch
        buf.copy_(t.from_numpy(host))
        if self.world > 1:
            self._dist.all_reduce(buf, op=self._dist.ReduceOp.SUM, group=self.group)
        return buf.cpu().numpy()

    def evaluate_gradient(self, G, pulsevals):
        e = self.engine
        if self._pipe is not None and self._pipe.exchange == "p2p":
            # peers attached: the blocking C-ABI call IS the sharded evaluation (one graph launch per rank, the
            # exchanges happen inside its kernels); every rank gets the global J and gradient
            J = e.evaluate_gradient(G, pulsevals)
            self.J_parts[:] = e.J_parts
            self.grad_J_Tb[:] = e.grad_J_Tb
            self.grad_J_a[:] = e.grad_J_a
            return J
        if self._pipe is not None:
            t = self._torch
            LNT = G.shape[0]
            self._h_eps.numpy()[:] = pulsevals
            with t.cuda.stream(self._stream):
                self._d_eps.copy_(self._h_eps, non_blocking=True)
                self._pipe.step(self._d_eps)
                self._h_out[:LNT].copy_(self._pipe.gradient(), non_blocking=True)
                self._h_out[LNT:].copy_(self._J_t, non_blocking=True)
            self._pipe.finish()          # one stream synchronisation + chi-norm / Taylor error flags
            out = self._h_out.numpy()
            G[:] = out[:LNT]
            self.J_parts[:] = out[LNT:]
            return float(np.sum(self.J_parts))
        sums = e.forward(pulsevals)
        sums_g = self._allreduce(self._sums, sums)
        self.J_parts[:] = e.backward(sums_g, self._G_partial)
        self.grad_J_Tb[:] = self._allreduce(self._grad, self._G_partial)
        self.grad_J_a[:] = e.grad_J_a
        G[:] = self.grad_J_Tb
        if self.problem.ja_kind:
            G += self.problem.lambda_a * self.grad_J_a
        return float(np.sum(self.J_parts))

    def evaluate_functional(self, pulsevals):
        if self._pipe is not None and self._pipe.exchange == "p2p":
            J = self.engine.evaluate_functional(pulsevals)
            self.J_parts[:] = self.engine.J_parts
            return J
        # forward only; J_T from the reduced sums (same formulas as finalize_J)
        p = self.problem
        sums = self._allreduce(self._sums, self.engine.forward(pulsevals))
        Kg = float(p.K_global)
        if p.functional == 0:
            JT = 1.0 - (sums[0] ** 2 + sums[1] ** 2) / Kg ** 2
        elif p.functional == 1:
            JT = 1.0 - sums[0] / Kg
        else:
            JT = 1.0 - sums[2] / Kg
        self.J_parts[0] = JT
        if p.ja_kind:
            dt = np.diff(p.tlist)
            e2 = np.asarray(pulsevals).reshape(p.L, p.NT) ** 2
            self.J_parts[1] = p.lambda_a * float(np.sum(e2 * dt[None, :]))
        self.J_parts[2] = p.lambda_b * sums[3] if p.gb_kind else 0.0
        return float(np.sum(self.J_parts))


class _DevArray:
    """Raw device pointer -> object torch.as_tensor understands."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(int(ptr), False), version=2)


class DevicePipeline:
    """Device-resident gradient evaluation: pulse values already in HBM, both
    exchanges done IN PLACE on the engine's device buffers by NCCL on the
    engine's stream, no host synchronisation inside a step.

    Call `step()` under `torch.cuda.stream(torch.cuda.ExternalStream(engine.stream()))`."""

    def __init__(self, engine, dist=None, group=None, exchange="p2p"):
        import torch
        self.engine, self.dist, self.group = engine, dist, group
        LNT = engine.L * engine.NT
        dev = torch.device("cuda", torch.cuda.current_device())
        self.sums_t = torch.as_tensor(_DevArray(engine.device_ptr(1), 4), device=dev)
        self.gTb_t = torch.as_tensor(_DevArray(engine.device_ptr(0), LNT), device=dev)
        self.G_t = torch.as_tensor(_DevArray(engine.device_ptr(3), LNT), device=dev)
        # J_T_sm: chi_k is proportional to sum_j tau_j (reference docs/src/tutorial.md:399-405), so the backward sweep
        # needs the global sums. J_T_re / J_T_ss: chi_k only depends on tau_k, the sums are needed for J alone, and
        # both exchanges travel together after the backward sweep.
        self.coupled = int(getattr(engine.problem, "functional", 0)) == 0
        self.coalesce = False
        self.exchange = "none"
        self.exchange_note = ""
        if dist is not None:
            self.exchange = exchange
            if exchange == "p2p" and not self._attach_peers(torch, dev):
                self.exchange = "nccl"
            if self.exchange == "nccl" and not self.coupled:
                self.coalesce = self._probe_coalescing(torch, dev)

    def _attach_peers(self, torch, dev):
        """Pass the IPC handles of the shards' exchange buffers around (the only use of torch.distributed) and map
        them. All ranks must end up on the same path: if any rank cannot map its peers, every rank falls back to NCCL
        (recorded in `exchange_note`, reported by bench.py)."""
        d = self.dist
        rank, world = d.get_rank(self.group), d.get_world_size(self.group)
        ok, note = 1.0, ""
        try:
            mine = self.engine.xchg_init(rank, world)
            handles = [None] * world
            d.all_gather_object(handles, mine, group=self.group)
            self.engine.xchg_attach(handles)
        except Exception as exc:          # GrapeError (no peer access / IPC refused) or a failed gather
            ok, note = 0.0, str(exc)
        flag = torch.tensor([ok], dtype=torch.float64, device=dev)
        d.all_reduce(flag, op=d.ReduceOp.MIN, group=self.group)
        if flag.item() != 1.0:
            try:
                self.engine.xchg_detach()
            except Exception:
                pass
            self.exchange_note = "p2p attach failed on some rank -> nccl" + (f" ({note})" if note else "")
            return False
        d.barrier(group=self.group)
        return True

    def _probe_coalescing(self, torch, dev):
        """one coalesced all-reduce of two scratch tensors on every rank: usable iff it returns the right sums"""
        d = self.dist
        try:
            a = torch.ones(4, dtype=torch.float64, device=dev)
            b = torch.full((8,), 2.0, dtype=torch.float64, device=dev)
            with d._coalescing_manager(group=self.group):
                d.all_reduce(a, op=d.ReduceOp.SUM, group=self.group)
                d.all_reduce(b, op=d.ReduceOp.SUM, group=self.group)
            w = float(d.get_world_size(self.group))
            ok = bool(torch.all(a == w).item() and torch.all(b == 2.0 * w).item())
        except Exception:
            ok = False
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        d.all_reduce(flag, op=d.ReduceOp.MIN, group=self.group)      # all ranks must take the same path
        return bool(flag.item() == 1.0)

    def step(self, d_eps):
        e, d = self.engine, self.dist
        if self.exchange == "p2p":        # exchanges inside the library's kernels (NVLink peer stores)
            e.enqueue_forward(d_eps.data_ptr())
            e.enqueue_backward()
            return
        e.enqueue_forward(d_eps.data_ptr())
        if d is not None and self.coupled:
            d.all_reduce(self.sums_t, op=d.ReduceOp.SUM, group=self.group)
        e.enqueue_backward()
        if d is not None:
            if self.coupled:
                d.all_reduce(self.gTb_t, op=d.ReduceOp.SUM, group=self.group)
            elif self.coalesce:
                with d._coalescing_manager(group=self.group):
                    d.all_reduce(self.sums_t, op=d.ReduceOp.SUM, group=self.group)
                    d.all_reduce(self.gTb_t, op=d.ReduceOp.SUM, group=self.group)
            else:
                d.all_reduce(self.sums_t, op=d.ReduceOp.SUM, group=self.group)
                d.all_reduce(self.gTb_t, op=d.ReduceOp.SUM, group=self.group)
            e.enqueue_combine()

    def finish(self):
        self.engine.finish()

    def close(self):
        """Unmap the peers' exchange buffers (after a barrier: no rank may still be pushing into ours)."""
        if self.exchange == "p2p" and self.dist is not None:
            self.engine.finish()
            self.dist.barrier(group=self.group)
            self.engine.xchg_detach()
            self.exchange = "none"

    def gradient(self):
        return self.G_t
