"""Loader for tests/golden/*.npz (see tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

from grape.jl_b200.problem import GrapeProblem

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    fn, gm, ja, gbk = (int(v) for v in z["scalars"])
    la, lb = (float(v) for v in z["lambdas"])
    p = GrapeProblem(z["tlist"], z["H0"], z["Hc"], z["psi0"], z["tgt"], gen_of_traj=z["gen_of_traj"],
                     shape=None if z["shape"].size == 0 else z["shape"],
                     weights=None if z["weights"].size == 0 else z["weights"],
                     functional=fn, gradient_method=gm, ja_kind=ja, lambda_a=la, gb_kind=gbk, lambda_b=lb,
                     gb_D=None if z["gb_D"].size == 0 else z["gb_D"], name=name)
    return p, z["pulsevals"], z
