"""GPU session helper (not a test): C3 shards of 4096/N trajectories on ONE GPU -- ms per gradient (device-resident
pipeline, CUDA events, L2 flushed) for several segment lengths S (GRAPE_B200_SEG_S) and both generations of the
real-symmetric kernels (GRAPE_B200_SYM_V).  Input for the strong-scaling segment-length rule."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402
from grape.jl_b200.engine import GrapeEngine  # noqa: E402
from grape.jl_b200.sharded import DevicePipeline  # noqa: E402


def measure(p, eps, steps=30, warm=5, **env):
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    try:
        e = GrapeEngine(p)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    pipe = DevicePipeline(e)
    dev = torch.device("cuda", 0)
    st = torch.cuda.ExternalStream(e.stream(), device=dev)
    d_eps = torch.from_numpy(eps).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(st):
        for _ in range(warm):
            pipe.step(d_eps)
    pipe.finish()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(st):
        for a, b in evs:
            flush.zero_()
            a.record(st)
            pipe.step(d_eps)
            b.record(st)
    pipe.finish()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    e.set_profiling(True)
    ph = np.zeros(7)
    for _ in range(10):
        e.eval_fg_device(d_eps.data_ptr(), None, None)
        t = e.timings()
        ph += np.array([t["formU_ms"], t["forward_ms"], t["tau_ms"], t["backward_ms"], t["gradient_ms"], t["d2h_ms"], t["total_ms"]])
    sched = e.small_schedule()
    e.close()
    return ms, (ph / 10).round(4).tolist(), sched


if __name__ == "__main__":
    out = []
    for nd in (64, 32, 16, 8):           # K = 4096, 2048, 1024, 512
        p, eps = configs.c3_ensemble(n_delta=nd, n_amp=64)
        for v in (2, 1):
            ms, ph, sched = measure(p, eps, GRAPE_B200_SYM_V=v, GRAPE_B200_FORCE_FORMSEG=1)
            out.append(dict(K=p.K, sym_v=v, S="auto", ms=ms, phases=ph, schedule=sched))
            print(json.dumps(out[-1]), flush=True)
        for S in (8, 10, 12, 16, 20, 25, 32, 40, 50):
            if nd == 64 and S < 20:
                continue
            ms, ph, sched = measure(p, eps, steps=20, GRAPE_B200_SEG_S=S, GRAPE_B200_FORCE_FORMSEG=1)
            out.append(dict(K=p.K, sym_v=2, S=S, ms=ms, phases=ph, schedule=sched))
            print(json.dumps(out[-1]), flush=True)
