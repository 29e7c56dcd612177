"""profiling helper (not a test): one gradient of a width-preserving reduced C5/C4."""
import sys
import numpy as np
sys.path.insert(0, ".")
from grape.jl_b200 import configs
from grape.jl_b200.engine import GrapeEngine
which, NT = sys.argv[1], int(sys.argv[2])
p, eps = (configs.c5_dense1024(NT=NT) if which == "c5" else configs.c4_dense450(NT=NT))
e = GrapeEngine(p)
G = np.zeros_like(eps)
for _ in range(2):
    J = e.evaluate_gradient(G, eps)
print(J, np.linalg.norm(G))
