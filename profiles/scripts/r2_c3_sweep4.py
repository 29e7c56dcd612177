"""GPU session helper (not a test): C3 full size on one GPU -- 128-register variant of the staged gradient kernel
(GRAPE_B200_SYM_OCC=4) and segment lengths around the automatic choice, with the final kernels."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r2_c3_sweep import measure  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402

if __name__ == "__main__":
    p, eps = configs.c3_ensemble()
    for label, env in (("auto", {}), ("occ4", dict(GRAPE_B200_SYM_OCC=4)), ("S=18", dict(GRAPE_B200_SEG_S=18)),
                       ("S=20", dict(GRAPE_B200_SEG_S=20)), ("S=22", dict(GRAPE_B200_SEG_S=22)), ("S=28", dict(GRAPE_B200_SEG_S=28)),
                       ("auto again", {})):
        ms, ph, sched = measure(p, eps, steps=40, **env)
        print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph[:5], schedule=sched)), flush=True)
