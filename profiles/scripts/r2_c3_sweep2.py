"""GPU session helper (not a test): C3 shards on ONE GPU with the scan schedule -- ms per gradient for
S in {8, 11, 16, 32} (NSEG ~ 125, 91, 63, 32) against the chain schedule (GRAPE_B200_SEG_SCAN=0)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r2_c3_sweep import measure  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402

if __name__ == "__main__":
    for nd in (64, 32, 16, 8):
        p, eps = configs.c3_ensemble(n_delta=nd, n_amp=64)
        ms, ph, sched = measure(p, eps)
        print(json.dumps(dict(K=p.K, mode="auto", ms=ms, phases=ph, schedule=sched)), flush=True)
        ms, ph, sched = measure(p, eps, GRAPE_B200_SEG_SCAN=0)
        print(json.dumps(dict(K=p.K, mode="chains,auto", ms=ms, phases=ph, schedule=sched)), flush=True)
        for S in (8, 11, 16, 21, 32):
            ms, ph, sched = measure(p, eps, steps=20, GRAPE_B200_SEG_S=S)
            print(json.dumps(dict(K=p.K, mode="scan", S=S, ms=ms, phases=ph, schedule=sched)), flush=True)
