#!/usr/bin/env python
"""bench.py -- GRAPE gradient evaluations/s (trajectory x time-steps per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = one complete gradient evaluation (forward sweep, chi, backward sweep,
gradient contraction, reduction) of the whole trajectory ensemble, i.e. one
`evaluate_gradient!` of the reference (src/optimize.jl:824-1014).

Workload at N=1 (default): BASELINE.json configs[2], the 4096-member three-level
robust ensemble (K=4096, N=3, L=2, NT=1000, J_T_ss) -- the configuration the
metric's "at 1/2/4/8 B200" is quoted on and the largest small-N config that is
not purely latency-bound.  configs[0]/[1] (500 / 8000 units per gradient) and the
dense configs[3]/[4] can be selected with --workload c1|c2|c4|c5.

Multi-GPU (torchrun, one rank per GPU): the ensemble is sharded over ranks
(weak scaling: every rank holds 4096 trajectories, the ensemble grows with N);
per gradient there are two all-reduces over NCCL (4 partial sums after the
forward sweep, the L*NT gradient after the backward sweep); for J_T_ss / J_T_re
(the default workload) they travel in one coalesced call after the backward sweep.

Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GRAPE gradient evals/sec (traj x timesteps/s)"
UNIT = "trajectory*timesteps/s"


def make_workload(name, world=1, scaling="weak"):
    from grape.jl_b200 import configs
    if name == "c3":
        n_delta = 64 * (world if scaling == "weak" else 1)
        p, eps = configs.c3_ensemble(n_delta=n_delta, n_amp=64)
        desc = dict(workload=f"c3_robust_ensemble K={p.K} N=3 L=2 NT=1000 J_T_ss (BASELINE configs[2])")
    elif name == "c1":
        p, eps = configs.c1_readme()
        desc = dict(workload="c1_readme_tls K=1 N=2 L=1 NT=500 J_T_sm (BASELINE configs[0])")
    elif name == "c2":
        p, eps = configs.c2_transmon()
        desc = dict(workload="c2_transmon_xgate K=4 N=6 L=2 NT=2000 J_T_sm (BASELINE configs[1])")
    elif name == "c4":
        kw = 16 * (world if scaling == "weak" else 1)      # weak scaling: 16 basis trajectories per GPU
        p, eps = configs.c4_dense450(K=kw)
        desc = dict(workload=f"c4_dense N=450 K={kw} L=2 NT=5000 J_T_sm (BASELINE configs[3])")
    elif name == "c5":
        kw = 64 * (world if scaling == "weak" else 1)      # weak scaling: 64 trajectories per GPU
        p, eps = configs.c5_dense1024(K=kw)
        desc = dict(workload=f"c5_dense N=1024 K={kw} L=2 NT=1000 J_T_sm+J_a+g_b (BASELINE configs[4])")
    else:
        raise SystemExit(f"unknown workload {name}")
    desc.update(K=p.K, N=p.N, L=p.L, NT=p.NT, units_per_step=p.K * p.NT)
    return p, eps, desc


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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
def cpu_reference_run(p, eps, sample_k, sample_nt, steps, warmup):
    """Times the C restatement of the reference algorithm (oracle/grape_oracle_c.c)
    on all host threads over a bounded sample of the workload."""
    from oracle import c_oracle as co
    # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1, so the count is passed explicitly
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or co.max_threads()
    sample_k = min(sample_k, p.K)
    sample_nt = min(sample_nt, p.NT)
    if p.N <= 32:   # dense configs: one sample already takes minutes of core time, no separate warm-up
        for _ in range(warmup):
            co.evaluate_gradient(p, eps, k_count=min(sample_k, 4 * cores), nt_count=min(sample_nt, 50), nthreads=cores)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        co.evaluate_gradient(p, eps, k_count=sample_k, nt_count=sample_nt, nthreads=cores)
        ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts))
    return dict(value=sample_k * sample_nt / t, unit=UNIT, cores=cores, kind="port",
                sample=f"{sample_k} of {p.K} trajectories x {sample_nt} of {p.NT} time steps per step, "
                       f"{steps} steps; C restatement of GRAPE.jl ExpProp+GradGenerator "
                       f"(dense Pade expm of N and N(L+1) matrices per trajectory-step), OpenMP over trajectories"), t


def cpu_sample_size(name):
    # sized for roughly 10-30 core-seconds of CPU work
    # c4: ~30 core-seconds per unit (1350 x 1350 block exponential); c5: ~5 core-minutes per unit (3072 x 3072),
    # so its sample is one trajectory-step per host thread, once
    return {"c1": (1, 500), "c2": (4, 2000), "c3": (1024, 1000), "c4": (16, 1), "c5": (16, 1)}[name]


# ---------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        p, eps, desc = make_workload(args.workload, 1)
        sk, snt = cpu_sample_size(args.workload)
        steps = max(1, min(args.steps, 5))
        cb, t = cpu_reference_run(p, eps, sk, snt, steps, min(args.warmup, 1))
        line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps,
                    warmup=min(args.warmup, 1), ms_per_step=t * 1e3, higher_is_better=True,
                    scaling=args.scaling, vs_baseline=None, dtype="f64 (complex128)", data="synthetic",
                    config=desc, impl="reference", cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from grape.jl_b200.engine import GrapeEngine
    from grape.jl_b200.sharded import DevicePipeline

    p, eps, desc = make_workload(args.workload, world, args.scaling)
    local = p.shard(rank, world) if world > 1 else p
    eng = GrapeEngine(local, device=local_rank)
    pipe = DevicePipeline(eng, dist if world > 1 else None)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    LNT = p.L * p.NT
    units_per_step = p.K * p.NT
    G = np.zeros(LNT)

    # ---- correctness guard: device-resident pipeline == host API result
    Jh = eng.evaluate_gradient(G, eps) if world == 1 else None
    d_eps = torch.from_numpy(eps).cuda()
    with torch.cuda.stream(stream):
        pipe.step(d_eps)
    pipe.finish()
    if world == 1:
        Gd = pipe.gradient().cpu().numpy()
        assert np.array_equal(Gd, G), "device pipeline and host API disagree"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`): one CUDA-event pair per step on the engine's stream, L2 flushed
    # ---- (256 MB device memset, outside the event pairs) before every timed step
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush_l2():
        with torch.cuda.stream(stream):
            flush_buf.zero_()

    for _ in range(args.warmup):
        with torch.cuda.stream(stream):
            pipe.step(d_eps)
    pipe.finish()
    eng.set_profiling(False)
    sampler = ClockSampler(local_rank)
    phase = np.zeros(8)
    barrier()
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launch_count()
    with torch.cuda.stream(stream):
        for e0, e1 in evs:
            flush_buf.zero_()
            e0.record(stream)
            pipe.step(d_eps)
            e1.record(stream)
    pipe.finish()
    barrier()
    launches = eng.launch_count() - launches0
    ms_total = float(sum(e0.elapsed_time(e1) for e0, e1 in evs))
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- per-phase kernel durations (CUDA events inside the library, same stream)
    roof = None
    if world == 1:
        n_prof = min(args.steps, 50)
        eng.set_profiling(True)
        for _ in range(n_prof):
            flush_l2()
            eng.eval_fg_device(d_eps.data_ptr(), None, None)
            tm = eng.timings()
            phase += np.array([tm["formU_ms"], tm["forward_ms"], tm["tau_ms"], tm["backward_ms"],
                               tm["gradient_ms"], tm["d2h_ms"], tm["total_ms"], tm["launches"]])
        phase /= n_prof
        eng.set_profiling(False)

    # ---- end-to-end through the public host API (`e2e`): host buffers, H2D + D2H inside
    for _ in range(3):
        eng.evaluate_gradient(G, eps) if world == 1 else None
    e2e = None
    if world == 1:
        x = eps.copy()
        barrier()
        t_sum = 0.0
        for i in range(args.steps):
            x[0] = eps[0] + 1e-9 * i          # new pulse values every step
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.evaluate_gradient(G, x)       # blocking: H2D pulses, kernels, D2H gradient, stream sync
            t_sum += time.perf_counter() - t0
        barrier()
        t_e2e = t_sum / args.steps
        d2h = 8 * (3 * LNT + 3 + 4 + 1 + 2 * p.K + 4)
        e2e = dict(value=units_per_step / t_e2e, unit=UNIT, ms_per_step=t_e2e * 1e3,
                   h2d_bytes_per_step=8 * LNT, d2h_bytes_per_step=d2h)
    else:
        # host-API path over ranks: numpy in, numpy out, collectives on staged device tensors
        from grape.jl_b200.sharded import ShardedGrape
        sh = ShardedGrape(p, lambda lp: eng, rank=rank, world=world, device=torch.device("cuda", local_rank))
        for _ in range(3):
            sh.evaluate_gradient(G, eps)
        x = eps.copy()
        barrier()
        n_e2e = max(10, args.steps // 4)
        t_sum = 0.0
        for i in range(n_e2e):
            x[0] = eps[0] + 1e-9 * i
            flush_l2()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            sh.evaluate_gradient(G, x)
            t_sum += time.perf_counter() - t0
        barrier()
        t = torch.tensor([t_sum / n_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
        e2e = dict(value=units_per_step / t_e2e, unit=UNIT, ms_per_step=t_e2e * 1e3,
                   h2d_bytes_per_step=8 * (LNT + 4), d2h_bytes_per_step=8 * (2 * LNT + 3 + 4 + 2 * local.K + 8))
        sh.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "of measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "of fallback 6650 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype="f64 (complex128)", data="synthetic",
                config=dict(desc, l2="flushed before every timed step by a 256 MB device memset issued outside the "
                                     "timed CUDA-event pair (value) / wall-clock window (e2e)",
                            parallelism=f"trajectory-sharded x{world}" if world > 1 else "single GPU"),
                e2e=e2e, gpu_launches=int(launches), clocks=clocks)
    if world == 1:
        from grape.jl_b200 import peaks as pk
        names = ["propagator_formation", "forward_sweep", "tau", "backward_sweep", "gradient_contraction"]
        dom = int(np.argmax(phase[:5]))
        k_ms = float(phase[dom])
        fp = pk.measure(local_rank)
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
            if tj and tj.get("phase") == names[dom]:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        if p.N > 32:
            # dense path: FP64 tensor-core (DMMA) roofline; peak measured on this GPU in this run
            form = eng.gradient_form()
            f_fwd, f_bwd, f_con, terms = pk.dense_flops_per_unit(p, eps, form)
            f_step = f_fwd + f_bwd + f_con
            share = {1: f_fwd, 3: f_bwd, 4: f_con}.get(dom, 0.0)
            kname = {1: "forward_sweep (dense_chain / dense2_chain, DMMA m8n8k4)",
                     3: "backward_sweep (" + ("dense_chain / dense2_chain" if form else "dense_backward / dense2_backward")
                        + ", DMMA m8n8k4)",
                     4: "gradient_contraction (kry_contract, DMMA m8n8k4)"}.get(dom, names[dom])
            achieved = share * units_per_step / (k_ms * 1e-3) / 1e12
            line["roofline"] = dict(
                bound="tensor", kernel=kname,
                achieved=achieved, peak=fp["dmma_tflops"], unit="TFLOP/s",
                frac=achieved / fp["dmma_tflops"] if fp["dmma_tflops"] > 0 else None, traffic=traffic,
                peak_source="FP64 DMMA peak measured on this GPU in this run (csrc/peaks.cu); MEASURED_PEAKS.json "
                            "carries no FP64 figure",
                kernel_ms=k_ms, flops_per_unit_kernel=share, flops_per_unit_step=f_step, taylor_terms=terms,
                gradient_form={0: "block_recursion", 1: "krylov", 2: "krylov, 2-3 Taylor terms per grid barrier"}.get(form, str(form)),
                flops_per_unit_reference_count=8.0 * p.N * p.N * terms * (2 + 2 * p.L),
                step_tflops=f_step * units_per_step / (ms_per_step * 1e-3) / 1e12,
                step_frac=f_step * units_per_step / (ms_per_step * 1e-3) / 1e12 / fp["dmma_tflops"],
                phase_tflops={nm: (fl * units_per_step / (float(ph) * 1e-3) / 1e12 if ph > 0 else None)
                              for nm, fl, ph in (("forward_sweep", f_fwd, phase[1]), ("backward_sweep", f_bwd, phase[3]),
                                                 ("gradient_contraction", f_con, phase[4]))},
                phase_ms=dict(zip(names, [float(v) for v in phase[:5]])),
                share_of_step=k_ms / float(phase[6]) if phase[6] > 0 else None)
        else:
            b_unit = 32 * p.N + 16 * p.L            # SURVEY 8d algorithmic bytes per unit
            achieved = b_unit * units_per_step / (k_ms * 1e-3) / 1e9
            line["roofline"] = dict(bound="hbm", kernel=names[dom],
                                    note="contract figure on ALGORITHMIC bytes (32N+16L per unit); the small-N kernels are "
                                         "FP64-FMA bound and, for Hermitian generators, recompute the forward states "
                                         "instead of re-reading them: roofline_fp64 is the binding roofline",
                                    achieved=achieved, peak=hbm_peak, unit="GB/s",
                                    frac=achieved / hbm_peak, traffic=traffic, traffic_source=traffic_src,
                                    peak_source=peak_src, kernel_ms=k_ms, algorithmic_bytes_per_unit=b_unit,
                                    phase_ms=dict(zip(names, [float(v) for v in phase[:5]])),
                                    share_of_step=k_ms / float(phase[6]) if phase[6] > 0 else None)
            sched = eng.small_schedule()
            fl_unit = pk.executed_flops_per_unit(p, eps, schedule=sched)
            line["roofline_fp64"] = dict(
                bound="fp64_fma", note="the small-N kernels are FP64-FMA bound, not HBM bound (SURVEY 8d); "
                                       "flops are the FP64 work the kernels execute (model in grape.jl_b200/peaks.py), step-level",
                achieved=fl_unit * units_per_step / (ms_per_step * 1e-3) / 1e12, unit="TFLOP/s",
                peak=fp["dfma_tflops"], peak_source="measured on this GPU in this run (csrc/peaks.cu DFMA loop)",
                dmma_peak=fp["dmma_tflops"], flops_per_unit=fl_unit,
                small_schedule={0: "n/a", 1: "segmented, general generators", 2: "segmented, Hermitian generators",
                                3: "segmented, real-symmetric generators"}.get(sched, str(sched)),
                frac=fl_unit * units_per_step / (ms_per_step * 1e-3) / 1e12 / fp["dfma_tflops"] if fp["dfma_tflops"] > 0 else None)
        if not args.no_cpu_baseline:
            sk, snt = cpu_sample_size(args.workload)
            cb, _ = cpu_reference_run(p, eps, sk, snt, 1, 1)
            line["cpu_baseline"] = cb
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
