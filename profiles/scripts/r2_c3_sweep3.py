"""GPU session helper (not a test): C3 shards on ONE GPU, final schedule rules -- auto choice per shard size against
forced alternatives (scan on/off, one-pass chains on/off)."""
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
        for label, env in (("auto", {}), ("scan=1", dict(GRAPE_B200_SEG_SCAN=1)), ("scan=0", dict(GRAPE_B200_SEG_SCAN=0)),
                           ("scan=0,dual=0", dict(GRAPE_B200_SEG_SCAN=0, GRAPE_B200_CHAIN_DUAL=0))):
            ms, ph, sched = measure(p, eps, **env)
            print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph, schedule=sched)), flush=True)
        for S in ((16, 20, 25, 32) if nd == 64 else (8, 11, 16)):
            ms, ph, sched = measure(p, eps, steps=20, GRAPE_B200_SEG_S=S)
            print(json.dumps(dict(K=p.K, mode="auto", S=S, ms=ms, phases=ph, schedule=sched)), flush=True)
