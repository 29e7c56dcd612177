"""timing helper (not a test): C3 gradient time against the segment length (GRAPE_B200_SEG_S; 0 = automatic)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from grape.jl_b200 import configs
from grape.jl_b200.engine import GrapeEngine
p, eps = configs.c3_ensemble()
ref = None
for S in [int(a) for a in sys.argv[1:]] or [0]:
    if S:
        os.environ["GRAPE_B200_SEG_S"] = str(S)
    else:
        os.environ.pop("GRAPE_B200_SEG_S", None)
    e = GrapeEngine(p)
    e.set_profiling(True)
    G = np.zeros_like(eps)
    for i in range(5):
        J = e.evaluate_gradient(G, eps)
    if ref is None:
        ref = G.copy()
    ts = []
    tm = []
    for i in range(30):
        t = time.perf_counter(); J = e.evaluate_gradient(G, eps); ts.append(time.perf_counter() - t)
        tm.append(e.timings())
    med = {k: float(np.median([t[k] for t in tm])) for k in tm[0]}
    print("S", S, "J", J, "dG", float(np.max(np.abs(G - ref))), "wall ms %.4f" % (np.median(ts) * 1e3),
          "units/s %.4g" % (p.K * p.NT / np.median(ts)), {k: round(v, 4) for k, v in med.items()})
    e.close()
