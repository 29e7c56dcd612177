"""GPU session helper (not a test): shards of the C3 ensemble on one GPU -- two-warp blocks of the gradient kernel
(automatic for under-filled launches) against four-warp blocks (GRAPE_B200_SYM_BD=128)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r2_c3_sweep import measure  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402

if __name__ == "__main__":
    for nd in (64, 32, 16, 8, 4):
        p, eps = configs.c3_ensemble(n_delta=nd, n_amp=64)
        for label, env in (("auto", {}), ("bd128", dict(GRAPE_B200_SYM_BD=128))):
            ms, ph, sched = measure(p, eps, steps=40, **env)
            print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph[:5], schedule=sched)), flush=True)
