"""GPU session helper (not a test): C3 full size, A/B of the pulse-value prefetch in the gradient kernel
(GRAPE_B200_SYM_OCC=5 = without), interleaved repetitions to beat run-to-run noise."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from r2_c3_sweep import measure  # noqa: E402
from grape.jl_b200 import configs  # noqa: E402

if __name__ == "__main__":
    p, eps = configs.c3_ensemble()
    for rep in range(3):
        for label, env in (("prefetch", {}), ("no prefetch", dict(GRAPE_B200_SYM_OCC=5))):
            ms, ph, sched = measure(p, eps, steps=60, **env)
            print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph[:5], schedule=sched)), flush=True)
