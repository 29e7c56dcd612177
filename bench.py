#!/usr/bin/env python
"""bench.py -- GRAPE gradient evaluations/s (trajectory x time-steps per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = one complete gradient evaluation (forward sweep, chi, backward sweep,
gradient contraction, reduction over the ensemble) of the whole trajectory ensemble,
i.e. one `evaluate_gradient!` of the reference (src/optimize.jl:824-1014).

Workload (default): BASELINE.json configs[2] AS WRITTEN -- the 4096-member three-level
robust ensemble (K=4096, N=3, L=2, NT=1000, J_T_ss), "sharded over 1/2/4/8 GPUs":
at N GPUs every rank holds 4096/N trajectories ("scaling": "strong").  The weak-scaling
variant (4096 trajectories PER GPU) is measured in the same run and reported under the
extra key `weak`.  configs[0]/[1] (500 / 8000 units per gradient) and the dense
configs[3]/[4] can be selected with --workload c1|c2|c4|c5; at N=1 short runs of c4 and
c5 (the FP64 tensor-core path) are appended as `extra_workloads`.

Multi-GPU (torchrun, one rank per GPU): trajectories are sharded over the ranks; the two
couplings (4 partial sums of tau, the L*NT partial gradient) are reduced by the library's
own kernels over NVLink peer memory (csrc/xchg.cuh) -- torch.distributed only passes the
CUDA IPC handles around at set-up (and provides the barrier / max-over-ranks of the timing).
Before timing, every N>1 run checks the sharded gradient against the 1-GPU gradient of the
same ensemble (1e-12), that it is bit-identical on every rank and run to run.

Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GRAPE gradient evals/sec (traj x timesteps/s)"
UNIT = "trajectory*timesteps/s"
DTYPE = "f64 (complex128)"
L2_NOTE = ("flushed before every timed step by a 256 MB device memset issued outside the timed CUDA-event pair "
           "(value) / wall-clock window (e2e)")


def make_workload(name, world=1, scaling="strong"):
    from grape.jl_b200 import configs
    mult = world if scaling == "weak" else 1
    if name == "c3":
        p, eps = configs.c3_ensemble(n_delta=64 * mult, n_amp=64)
        wl = f"c3_robust_ensemble K={p.K} N=3 L=2 NT=1000 J_T_ss (BASELINE configs[2])"
    elif name == "c1":
        p, eps = configs.c1_readme()
        wl = "c1_readme_tls K=1 N=2 L=1 NT=500 J_T_sm (BASELINE configs[0])"
    elif name == "c2":
        p, eps = configs.c2_transmon()
        wl = "c2_transmon_xgate K=4 N=6 L=2 NT=2000 J_T_sm (BASELINE configs[1])"
    elif name == "c4":
        p, eps = configs.c4_dense450(K=16 * mult)
        wl = f"c4_dense N=450 K={p.K} L=2 NT=5000 J_T_sm (BASELINE configs[3])"
    elif name == "c5":
        p, eps = configs.c5_dense1024(K=64 * mult)
        wl = f"c5_dense N=1024 K={p.K} L=2 NT=1000 J_T_sm+J_a+g_b (BASELINE configs[4])"
    else:
        raise SystemExit(f"unknown workload {name}")
    # ONE config dict for both arms (GPU and --impl reference): the driver compares them
    cfg = dict(workload=wl, K=p.K, N=p.N, L=p.L, NT=p.NT, units_per_step=p.K * p.NT, l2=L2_NOTE,
               parallelism=(f"{p.K} trajectories sharded over {world} GPUs ({p.K // world} per GPU), "
                            f"sums and gradient reduced over NVLink peer memory inside the library"
                            if world > 1 else "single GPU"))
    return p, eps, cfg


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------- CPU arm
def host_cores():
    # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1, so the count is passed explicitly
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


CPU_KIND_NOTE = ("C restatement of GRAPE.jl ExpProp+GradGenerator (dense Pade expm of the N x N and N(L+1) x N(L+1) "
                 "matrices per trajectory-step, no sharing between trajectories), OpenMP over trajectories like the "
                 "reference's @threadsif (src/optimize.jl:720, 876)")


def cpu_reference_run(p, eps, steps, warmup, sample_k=None, sample_nt=None):
    """Times oracle/grape_oracle_c.c (the reference algorithm as the reference executes it) on all host threads.
    Small-N workloads: the FULL ensemble, every step. Dense workloads (one trajectory-step = 30 core-seconds at
    N=450, 5 core-minutes at N=1024): a bounded sample, cost exactly linear in K and NT."""
    from oracle import c_oracle as co
    cores = host_cores()
    sk = p.K if sample_k is None else min(sample_k, p.K)
    snt = p.NT if sample_nt is None else min(sample_nt, p.NT)
    for _ in range(warmup):
        co.evaluate_gradient(p, eps, k_count=sk, nt_count=snt, nthreads=cores)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        co.evaluate_gradient(p, eps, k_count=sk, nt_count=snt, nthreads=cores)
        ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    full = sk == p.K and snt == p.NT
    sample = (f"the full workload ({p.K} trajectories x {p.NT} time steps) per step" if full else
              f"{sk} of {p.K} trajectories x {snt} of {p.NT} time steps per step (cost is linear in both)")
    return dict(value=sk * snt / t, unit=UNIT, cores=cores, kind="port",
                sample=f"{sample}, {steps} timed steps after {warmup} warm-up; {CPU_KIND_NOTE}"), t


def cpu_sample_size(name):
    # dense configs only: one trajectory-step per host thread (c4: ~30 core-seconds each, c5: ~5 core-minutes)
    return {"c4": (16, 1), "c5": (16, 1)}.get(name, (None, None))


# ---------------------------------------------------------------------------- GPU measurement
class GpuRun:
    """One engine (+ peers) and the timing helpers shared by the headline workload and the extra workloads."""

    def __init__(self, p, eps, rank, local_rank, world, dist, exchange):
        import torch
        from grape.jl_b200.engine import GrapeEngine
        from grape.jl_b200.sharded import DevicePipeline
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.p, self.eps = p, eps
        self.local = p.shard(rank, world) if world > 1 else p
        self.dev = torch.device("cuda", local_rank)
        self.eng = GrapeEngine(self.local, device=local_rank)
        self.pipe = DevicePipeline(self.eng, dist if world > 1 else None, exchange=exchange)
        self.stream = torch.cuda.ExternalStream(self.eng.stream(), device=self.dev)
        self.d_eps = torch.from_numpy(eps).to(self.dev)
        self.LNT = p.L * p.NT
        self.units = p.K * p.NT
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        # the reference-facing blocking call: with peers attached (or one GPU) that is the C-ABI eval_fg itself
        self.sh = None
        self.host_eval = self.eng.evaluate_gradient
        if world > 1 and self.pipe.exchange == "nccl":
            from grape.jl_b200.sharded import ShardedGrape
            self.sh = ShardedGrape(p, lambda lp: self.eng, rank=rank, world=world, device=self.dev, exchange="nccl")
            self.host_eval = self.sh.evaluate_gradient

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def device_gradient(self):
        with self.torch.cuda.stream(self.stream):
            self.pipe.step(self.d_eps)
        self.pipe.finish()
        return self.pipe.gradient().cpu().numpy().copy()

    def time_device_steps(self, steps, warmup):
        """`value`: pulse values resident in HBM, one CUDA-event pair per step on the engine's stream, L2 flushed
        (outside the pairs) before every step; max over ranks of the summed durations."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.pipe.step(self.d_eps)
        self.pipe.finish()
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = self.eng.launch_count()
        with torch.cuda.stream(self.stream):
            for e0, e1 in evs:
                self.flush_buf.zero_()
                e0.record(self.stream)
                self.pipe.step(self.d_eps)
                e1.record(self.stream)
        self.pipe.finish()
        self.barrier()
        launches = self.eng.launch_count() - l0
        ms = float(sum(e0.elapsed_time(e1) for e0, e1 in evs))
        return self.max_over_ranks(ms) / steps, launches

    def time_host_api(self, steps, warmup=3):
        """`e2e`: the blocking reference-facing call with HOST buffers -- pulse H2D, kernels (and exchanges),
        gradient D2H, stream synchronisation -- timed by wall clock; new pulse values every step."""
        torch = self.torch
        G = np.zeros(self.LNT)
        x = self.eps.copy()
        for _ in range(warmup):
            self.host_eval(G, x)
        self.barrier()
        t_sum = 0.0
        for i in range(steps):
            x[0] = self.eps[0] + 1e-9 * i
            with torch.cuda.stream(self.stream):
                self.flush_buf.zero_()
            torch.cuda.synchronize()
            if self.dist is not None:
                self.dist.barrier()
            t0 = time.perf_counter()
            self.host_eval(G, x)
            t_sum += time.perf_counter() - t0
        self.barrier()
        t = self.max_over_ranks(t_sum / steps)
        K = self.local.K
        return dict(value=self.units / t, unit=UNIT, ms_per_step=t * 1e3, h2d_bytes_per_step=8 * self.LNT,
                    d2h_bytes_per_step=8 * (3 * self.LNT + 3 + 4 + 1 + 2 * K + 6)), G

    def phases(self, n_prof):
        """per-phase kernel durations of one evaluation (CUDA events inside the library, same stream)"""
        phase = np.zeros(8)
        self.eng.set_profiling(True)
        for _ in range(n_prof):
            with self.torch.cuda.stream(self.stream):
                self.flush_buf.zero_()
            self.eng.eval_fg_device(self.d_eps.data_ptr(), None, None)
            tm = self.eng.timings()
            phase += np.array([tm["formU_ms"], tm["forward_ms"], tm["tau_ms"], tm["backward_ms"],
                               tm["gradient_ms"], tm["d2h_ms"], tm["total_ms"], tm["launches"]])
        self.eng.set_profiling(False)
        return phase / n_prof

    def close(self):
        self.pipe.finish()
        self.barrier()
        if self.sh is not None:
            self.sh.close()
        self.pipe.close()
        self.flush_buf = None
        self.eng.close()


PHASES = ["propagator_formation", "forward_sweep", "tau", "backward_sweep", "gradient_contraction"]


def roofline_small(run, phase, ms_per_step, fp, hbm_peak, peak_src, workload):
    """N <= 32: the kernels are FP64-FMA bound; `roofline` = the dominant kernel's executed FP64 flops
    over its own duration against the DFMA peak measured in this run; `roofline_hbm` keeps the algorithmic-bytes
    figure (32N + 16L per unit) for the same kernel."""
    from grape.jl_b200 import peaks as pk
    p, eps = run.p, run.eps
    dom = int(np.argmax(phase[:5]))
    k_ms = float(phase[dom])
    sched = run.eng.small_schedule()
    fl = pk.executed_flops_split(p, eps, schedule=sched)
    by_phase = {0: fl["formation"], 1: fl["chains"] / 2, 3: fl["chains"] / 2, 4: fl["contraction"]}
    fl_step = fl["formation"] + fl["chains"] + fl["contraction"]
    share = by_phase.get(dom, 0.0)
    ach = share * run.units / (k_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if tj and tj.get("phase") == PHASES[dom]:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    sched_name = {0: "n/a", 1: "segmented, general generators", 2: "segmented, Hermitian generators",
                  3: "segmented, real-symmetric generators"}.get(sched, str(sched))
    roof = dict(bound="fp64_fma", kernel=PHASES[dom], achieved=ach, peak=fp["dfma_tflops"], unit="TFLOP/s",
                frac=ach / fp["dfma_tflops"] if fp["dfma_tflops"] > 0 else None, traffic=traffic,
                traffic_source=traffic_src,
                peak_source="FP64 FMA (DFMA) peak measured on this GPU in this run (csrc/peaks.cu); MEASURED_PEAKS.json "
                            "carries no FP64 figure; cuBLAS ZGEMM 8192^3 beside it in `peaks`",
                kernel_ms=k_ms, flops_per_unit_kernel=share, flops_per_unit_step=fl_step, small_schedule=sched_name,
                note="executed FP64 flops (model in grape.jl_b200/peaks.py, cross-checked against ncu DFMA counts in "
                     "round 1); the binding bound of the small-N kernels is FP64 issue, not HBM (SURVEY 8d)",
                step_tflops=fl_step * run.units / (ms_per_step * 1e-3) / 1e12,
                step_frac=fl_step * run.units / (ms_per_step * 1e-3) / 1e12 / fp["dfma_tflops"] if fp["dfma_tflops"] > 0 else None,
                phase_ms=dict(zip(PHASES, [float(v) for v in phase[:5]])),
                share_of_step=k_ms / float(phase[6]) if phase[6] > 0 else None)
    b_unit = 32 * p.N + 16 * p.L            # SURVEY 8d algorithmic bytes per unit
    ach_b = b_unit * run.units / (k_ms * 1e-3) / 1e9
    roof_hbm = dict(bound="hbm", kernel=PHASES[dom], achieved=ach_b, peak=hbm_peak, unit="GB/s", frac=ach_b / hbm_peak,
                    traffic=traffic, peak_source=peak_src, algorithmic_bytes_per_unit=b_unit,
                    note="secondary: ALGORITHMIC bytes over the kernel time; for Hermitian generators the kernels recompute "
                         "the forward states instead of re-reading them, so DRAM traffic is far below this figure")
    return roof, roof_hbm


def roofline_dense(run, phase, ms_per_step, fp, workload):
    from grape.jl_b200 import peaks as pk
    p, eps = run.p, run.eps
    dom = int(np.argmax(phase[:5]))
    k_ms = float(phase[dom])
    form = run.eng.gradient_form()
    f_fwd, f_bwd, f_con, terms = pk.dense_flops_per_unit(p, eps, form)
    f_step = f_fwd + f_bwd + f_con
    share = {1: f_fwd, 3: f_bwd, 4: f_con}.get(dom, 0.0)
    if getattr(run.eng, "dense_concurrent", lambda: False)() and dom in (1, 3):   # both chains share one kernel sequence
        share = f_fwd + f_bwd
    kname = {1: "forward_sweep (dense_chain / dense2_chain, DMMA m8n8k4)",
             3: "backward_sweep (" + ("dense_chain / dense2_chain" if form else "dense_backward / dense2_backward")
                + ", DMMA m8n8k4)",
             4: "gradient_contraction (kry_contract, DMMA m8n8k4)"}.get(dom, PHASES[dom])
    ach = share * run.units / (k_ms * 1e-3) / 1e12
    return dict(
        bound="tensor", kernel=kname, achieved=ach, peak=fp["dmma_tflops"], unit="TFLOP/s",
        frac=ach / fp["dmma_tflops"] if fp["dmma_tflops"] > 0 else None, traffic=None,
        peak_source="FP64 DMMA (mma.sync.m8n8k4.f64) peak measured on this GPU in this run (csrc/peaks.cu); "
                    "MEASURED_PEAKS.json carries no FP64 figure; cuBLAS ZGEMM 8192^3 beside it in `peaks`",
        kernel_ms=k_ms, flops_per_unit_kernel=share, flops_per_unit_step=f_step, taylor_terms=terms,
        gradient_form={0: "block_recursion", 1: "krylov", 2: "krylov, 2-3 Taylor terms per grid barrier"}.get(form, str(form)),
        flops_per_unit_reference_count=8.0 * p.N * p.N * terms * (2 + 2 * p.L),
        step_tflops=f_step * run.units / (ms_per_step * 1e-3) / 1e12,
        step_frac=f_step * run.units / (ms_per_step * 1e-3) / 1e12 / fp["dmma_tflops"],
        phase_ms=dict(zip(PHASES, [float(v) for v in phase[:5]])),
        share_of_step=k_ms / float(phase[6]) if phase[6] > 0 else None)


def measure_peaks(local_rank):
    """FP64 denominators on this GPU, in this run: DFMA / DMMA loops (csrc/peaks.cu) and, as the library yardstick
    BASELINE.md asks for, cuBLAS ZGEMM 8192^3 through torch.matmul (8 N^3 flops)."""
    import torch
    from grape.jl_b200 import peaks as pk
    fp = pk.measure(local_rank)
    try:
        n = 8192
        a = torch.randn(n, n, dtype=torch.complex128, device=f"cuda:{local_rank}")
        b = torch.randn(n, n, dtype=torch.complex128, device=f"cuda:{local_rank}")
        torch.matmul(a, b)
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 8.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        fp["cublas_zgemm_8192_tflops"] = best
        del a, b
        torch.cuda.empty_cache()
    except Exception as exc:
        fp["cublas_zgemm_8192_tflops"] = None
        fp["cublas_zgemm_note"] = str(exc)[:200]
    return fp


def extra_workload(name, local_rank, fp, steps=2, warmup=1):
    """Short run of a dense config (the DMMA path) with its own value / e2e / roofline (N=1 only)."""
    p, eps, cfg = make_workload(name, 1)
    run = GpuRun(p, eps, 0, local_rank, 1, None, "p2p")
    ms, launches = run.time_device_steps(steps, warmup)
    e2e, _ = run.time_host_api(steps, warmup=1)
    phase = run.phases(1)
    roof = roofline_dense(run, phase, ms, fp, name)
    out = dict(config=cfg, value=run.units / (ms * 1e-3), unit=UNIT, ms_per_step=ms, steps=steps, warmup=warmup,
               e2e=e2e, gpu_launches=int(launches), roofline=roof)
    run.close()
    return out


# ---------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short c4 / c5 runs (extra_workloads)")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # the reference's own algorithm on the host cores (rank 0 only), same workload / config / steps as the GPU arm
        if rank != 0:
            return 0
        steps = 10 if args.steps is None else max(1, args.steps)
        warmup = 3 if args.warmup is None else max(0, args.warmup)
        p, eps, cfg = make_workload(args.workload, max(1, args.gpus), args.scaling)
        sk, snt = cpu_sample_size(args.workload)
        if sk is not None:
            warmup = 0      # dense: a single sample already takes minutes of core time
            steps = 1
        cb, t = cpu_reference_run(p, eps, steps, warmup, sk, snt)
        line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                    ms_per_step=t * 1e3, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype=DTYPE,
                    data="synthetic", config=cfg, impl="reference", cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    steps = 300 if args.steps is None else max(1, args.steps)
    warmup = max(3, 10 if args.warmup is None else args.warmup)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from grape.jl_b200.engine import GrapeEngine

    p, eps, cfg = make_workload(args.workload, world, args.scaling)
    run = GpuRun(p, eps, rank, local_rank, world, dist, args.exchange)
    LNT = run.LNT
    cfg["exchange"] = {"none": "n/a (single GPU)", "p2p": "in-library NVLink peer stores (csrc/xchg.cuh), no NCCL in the step",
                       "nccl": "NCCL all-reduce in place on the engine's buffers"}[run.pipe.exchange] + \
                      (f" [{run.pipe.exchange_note}]" if run.pipe.exchange_note else "")

    # ---- correctness guards (before any timing)
    guard = {}
    Gd = run.device_gradient()
    if world == 1:
        G = np.zeros(LNT)
        run.eng.evaluate_gradient(G, eps)
        assert np.array_equal(Gd, G), "device pipeline and host API disagree"
        guard = dict(device_pipeline_equals_host_api=True)
    else:
        # SURVEY gate G6: sharded result vs the 1-GPU result of the SAME ensemble, identical on every rank, run to run
        full = GrapeEngine(p, device=local_rank)
        G1 = np.zeros(LNT)
        J1 = full.evaluate_gradient(G1, eps)
        full.close()
        sc = float(np.max(np.abs(G1)))
        err = float(np.max(np.abs(Gd - G1))) / sc
        assert err <= 1e-12, f"sharded gradient differs from the 1-GPU gradient: rel {err:.3e}"
        Gd2 = run.device_gradient()
        assert np.array_equal(Gd, Gd2), "sharded gradient is not bit-identical run to run"
        t = torch.from_numpy(Gd.copy()).to(run.dev)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        assert same or run.pipe.exchange == "nccl", "sharded gradient differs between ranks"
        Gh = np.zeros(LNT)
        Jh = run.host_eval(Gh, eps)
        assert np.max(np.abs(Gh - G1)) <= 1e-12 * sc, "host-API sharded gradient differs from the 1-GPU gradient"
        assert abs(Jh - J1) <= 1e-12, "sharded J differs from the 1-GPU J"
        guard = dict(vs_single_gpu_rel_err=err, tolerance=1e-12, bit_identical_run_to_run=True,
                     identical_on_all_ranks=same, J_abs_err=float(abs(Jh - J1)))

    # ---- `value`: device-resident timing of exactly `steps` steps; clocks sampled under load (+ sustained window)
    sampler = ClockSampler(local_rank)
    run.barrier()
    if rank == 0:
        sampler.start()
    ms_per_step, launches = run.time_device_steps(steps, warmup)
    sustained = None
    if not args.no_sustained:
        n_sus = int(min(20000, max(steps, math.ceil(1000.0 / max(ms_per_step + 0.05, 1e-3)))))
        ms_sus, _ = run.time_device_steps(n_sus, 0)
        sustained = dict(steps=n_sus, ms_per_step=ms_sus, value=run.units / (ms_sus * 1e-3),
                         note="same measurement over a >= 1 s window (clock samples cover both windows)")
    clocks = sampler.stop() if rank == 0 else None
    value = run.units / (ms_per_step * 1e-3)

    # ---- per-phase kernel durations + end-to-end through the host API
    phase = run.phases(min(steps, 50)) if world == 1 else None
    e2e, _ = run.time_host_api(steps if world == 1 else max(10, steps // 2))

    # ---- weak-scaling variant (4096 trajectories per GPU), extra key
    weak = None
    if world > 1 and args.scaling == "strong" and args.workload in ("c3", "c4", "c5"):
        run.close()
        pw, epsw, cfgw = make_workload(args.workload, world, "weak")
        runw = GpuRun(pw, epsw, rank, local_rank, world, dist, args.exchange)
        runw.device_gradient()
        ms_w, _ = runw.time_device_steps(steps, warmup)
        e2e_w, _ = runw.time_host_api(max(10, steps // 2))
        weak = dict(scaling="weak", K_total=pw.K, value=runw.units / (ms_w * 1e-3), ms_per_step=ms_w, e2e=e2e_w,
                    workload=cfgw["workload"])
        runw.close()
        run = None

    if rank != 0:
        if run is not None:
            run.close()
        dist.destroy_process_group()
        return 0

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype=DTYPE, data="synthetic", config=cfg, e2e=e2e, gpu_launches=int(launches), clocks=clocks,
                parity_guard=guard)
    if sustained:
        line["sustained"] = sustained
    if weak:
        line["weak"] = weak
    if world == 1:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "of measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else \
            "of fallback 6650 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
        fp = measure_peaks(local_rank)
        line["peaks"] = fp
        if p.N > 32:
            line["roofline"] = roofline_dense(run, phase, ms_per_step, fp, args.workload)
        else:
            line["roofline"], line["roofline_hbm"] = roofline_small(run, phase, ms_per_step, fp, hbm_peak, peak_src,
                                                                   args.workload)
        run.close()
        run = None
        if not args.no_extra and args.workload == "c3":
            extras = {}
            for name in ("c4", "c5"):
                try:
                    extras[name] = extra_workload(name, local_rank, fp)
                except Exception as exc:      # an extra must never take the headline line down
                    extras[name] = dict(error=str(exc)[:300])
            line["extra_workloads"] = extras
        if not args.no_cpu_baseline:
            sk, snt = cpu_sample_size(args.workload)
            cb, _ = cpu_reference_run(p, eps, 3 if sk is None else 1, 1 if sk is None else 0, sk, snt)
            line["cpu_baseline"] = cb
    if run is not None:
        run.close()
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
