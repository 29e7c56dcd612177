"""GPU session helper (not a test): C3 full size, interleaved repetitions of the segment lengths around the rule."""
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
        for label, env in (("auto", {}), ("S=25", dict(GRAPE_B200_SEG_S=25)), ("S=20", dict(GRAPE_B200_SEG_S=20)),
                           ("S=23", dict(GRAPE_B200_SEG_S=23)), ("S=32", dict(GRAPE_B200_SEG_S=32))):
            ms, ph, sched = measure(p, eps, steps=60, **env)
            print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph[:5])), flush=True)
