import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from grape.jl_b200 import configs
from grape.jl_b200.engine import GrapeEngine
for name, (p, eps) in [("c1", configs.c1_readme()), ("c3", configs.c3_ensemble()), ("c3sm", configs.c3_ensemble(functional=0))]:
    e = GrapeEngine(p)
    e.set_profiling(True)
    G = np.zeros_like(eps)
    for i in range(3):
        J = e.evaluate_gradient(G, eps)
    ts = []
    for i in range(10):
        t = time.perf_counter(); J = e.evaluate_gradient(G, eps); ts.append(time.perf_counter() - t)
    print(name, "J", J, "|G|", np.linalg.norm(G), "wall ms", np.median(ts) * 1e3, "units/s", p.K * p.NT / np.median(ts))
    print("   ", e.timings())
