"""compute-sanitizer driver (not a test): one small gradient on every path / schedule."""
import sys
import numpy as np
sys.path.insert(0, ".")
import grape.jl_b200 as gb
from grape.jl_b200 import configs
from grape.jl_b200.engine import GrapeEngine
D3 = np.diag([0.0, 1.0, 0.0]).astype(complex)
cases = [
    configs.c1_readme(NT=37),
    configs.c3_ensemble(n_delta=5, n_amp=7, NT=41),
    configs.c3_ensemble(n_delta=3, n_amp=3, NT=20, gb_kind=1, gb_D=D3, lambda_b=0.3),          # plain chains
    configs.random_problem(K=5, N=4, L=3, NT=11, seed=1),
    configs.random_problem(K=3, N=3, L=2, NT=9, seed=2, gradient_method=gb.TAYLOR),
    configs.c2_transmon(NT=45),
    configs.random_problem(K=3, N=17, L=2, NT=7, seed=3, G=1),
    configs.random_problem(K=2, N=32, L=1, NT=5, seed=4, G=1),
    configs.c4_dense450(N=40, K=5, NT=3),
    configs.c5_dense1024(N=48, K=16, NT=3),
]
for p, eps in cases:
    e = GrapeEngine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    J2 = e.evaluate_functional(eps)
    e.stored_states(0)
    e.final_states()
    print(p.name, J, J2, float(np.linalg.norm(G)))
    e.close()
print("done")
