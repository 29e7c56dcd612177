"""compute-sanitizer driver (not a test): one small gradient on the schedules added in sessions 11-14 --
real-symmetric small path (fused formation forced), multi-term strip and tiled dense chains (3 / 2 terms, pre-formed and
in-kernel strips, Hermitian and non-Hermitian generators)."""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
import grape.jl_b200 as gb
from grape.jl_b200 import configs
from grape.jl_b200.engine import GrapeEngine


def run(p, eps, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        e = GrapeEngine(p)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    J = e.evaluate_gradient(G, eps)
    e.stored_states(0)
    e.final_states()
    print(p.name, env, J, float(np.linalg.norm(G)), e.small_schedule(), e.gradient_form())
    e.close()


p, eps = configs.c3_ensemble(n_delta=3, n_amp=5, NT=200)
run(p, eps, GRAPE_B200_FORCE_FORMSEG=1)
p, eps = configs.random_problem(K=7, N=2, L=3, NT=23, seed=5, real=True)
p.tlist[:] = p.tlist * 0.01
run(p, eps, GRAPE_B200_FORCE_FORMSEG=1)
for dense2 in (0, 1):
    for terms, pre in ((3, 1), (2, 1), (2, 0)):
        p, eps = configs.c4_dense450(N=40, K=16, NT=3)
        run(p, eps, GRAPE_B200_DENSE2=dense2, GRAPE_B200_DENSE_TERMS=terms, GRAPE_B200_DENSE_PREFORM=pre)
    p, eps = configs.random_problem(K=16, N=36, L=3, NT=3, G=1, seed=6, hermitian=False, shaped=True)
    p.tlist = p.tlist * (0.5 / 6.0)
    run(p, eps, GRAPE_B200_DENSE2=dense2)
    p, eps = configs.c5_dense1024(N=48, K=16, NT=3)
    run(p, eps, GRAPE_B200_DENSE2=dense2)
print("done")
