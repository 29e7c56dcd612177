"""GPU session helper (not a test): C3 shards (K = 2048, 1024, 512: the strong-scaling shards of 2, 4, 8 GPUs) with the
economised 128-register kernels -- segment length sweep of the scan schedule."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r2_c3_sweep import measure  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402

if __name__ == "__main__":
    for nd in (32, 16, 8):
        p, eps = configs.c3_ensemble(n_delta=nd, n_amp=64)
        for S in (None, 8, 9, 10, 11, 12, 13, 14, 16, 18, 20):
            env = dict(GRAPE_B200_SEG_S=S) if S else {}
            ms, ph, sched = measure(p, eps, steps=40, **env)
            print(json.dumps(dict(K=p.K, S=S, ms=ms, phases=ph[:5])), flush=True)
