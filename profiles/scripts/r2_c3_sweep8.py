"""GPU session helper (not a test): C3 full size and shards, economised polynomial (default) against the Taylor series
(GRAPE_B200_ECON=0), interleaved repetitions, and the segment length re-swept with the economised kernels."""
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
        for label, env in (("econ", {}), ("taylor", dict(GRAPE_B200_ECON=0))):
            ms, ph, sched = measure(p, eps, steps=60, **env)
            print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph[:5], sched=sched)), flush=True)
    for S in (16, 18, 20, 22, 25, 28, 32, 40):
        ms, ph, sched = measure(p, eps, steps=40, GRAPE_B200_SEG_S=S)
        print(json.dumps(dict(K=p.K, mode="econ S=%d" % S, ms=ms, phases=ph[:5])), flush=True)
    for nd in (32, 16, 8):
        p, eps = configs.c3_ensemble(n_delta=nd, n_amp=64)
        for label, env in (("econ", {}), ("taylor", dict(GRAPE_B200_ECON=0))):
            ms, ph, sched = measure(p, eps, steps=60, **env)
            print(json.dumps(dict(K=p.K, mode=label, ms=ms, phases=ph[:5], sched=sched)), flush=True)
    p, eps = configs.c1_readme()
    for label, env in (("econ", {}), ("taylor", dict(GRAPE_B200_ECON=0))):
        ms, ph, sched = measure(p, eps, steps=100, **env)
        print(json.dumps(dict(K=p.K, cfg="c1", mode=label, ms=ms, phases=ph[:5], sched=sched)), flush=True)
