"""GPU session helper (not a test): C3 full size, register / inline-order variants of the economised gradient kernel
(GRAPE_B200_SYM_OCC=6, 7, 8) against the default, interleaved, at several segment lengths."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r2_c3_sweep import measure  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402

if __name__ == "__main__":
    p, eps = configs.c3_ensemble()
    for rep in range(2):
        for occ in (None, 10):
            for S in (None, 25):
                env = {}
                if occ:
                    env["GRAPE_B200_SYM_OCC"] = occ
                if S:
                    env["GRAPE_B200_SEG_S"] = S
                ms, ph, sched = measure(p, eps, steps=40, **env)
                print(json.dumps(dict(K=p.K, occ=occ, S=S, ms=ms, phases=ph[:5], sched=sched)), flush=True)
