"""CPU: the C-ABI library builds, loads and exports every symbol the header
declares; host-side logic (configs, sharding, import shim, C oracle)."""
import ctypes
import os
import re

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs, _lib
from oracle import grape_oracle as go

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "grape_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(grape_b200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(lib, name), f"missing export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().grape_b200_abi_version() == 1


def test_descriptor_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "grape_b200.h")).read()
    body = re.search(r"typedef struct grape_b200_problem \{(.*?)\} grape_b200_problem;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([A-Za-z_0-9]+)\s*;", body)
    assert names == [f[0] for f in _lib.ProblemDesc._fields_]


def test_no_cpu_fallback_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from grape.jl_b200.engine import GrapeEngine, GrapeError
    with pytest.raises(GrapeError, match="no CPU fallback"):
        GrapeEngine(configs.c1_readme()[0])


def test_create_argument_validation(lib_built):
    lib = _lib.load()
    h = ctypes.c_void_p()
    d = _lib.ProblemDesc()
    d.abi_version = 99
    assert lib.grape_b200_create(ctypes.byref(d), ctypes.byref(h)) == 1
    d.abi_version = 1
    d.K, d.N, d.NT, d.G, d.L = 1, 2, 5, 1, 0
    assert lib.grape_b200_create(ctypes.byref(d), ctypes.byref(h)) == 5   # ENOCONTROLS
    assert b"no controls in trajectories" in lib.grape_b200_last_error(None)


def test_product_never_imports_oracle():
    """No file of the product (package, C-ABI sources, header, Julia shim) imports, links or executes oracle/."""
    roots = [os.path.join(ROOT, "grape.jl_b200"), os.path.join(ROOT, "grape"), os.path.join(ROOT, "include"),
             os.path.join(ROOT, "julia")]
    seen = 0
    for root in roots:
        for dirpath, _, files in os.walk(root):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                    continue
                src = open(os.path.join(dirpath, f)).read()
                seen += 1
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert not re.search(r"importlib[^\n]*oracle|libgrape_oracle|grape_oracle_c|dense_oracle", src), f
    assert seen > 20


def test_configs_shapes():
    for name, (K, N, L, NT) in {"c1": (1, 2, 1, 500), "c2": (4, 6, 2, 2000)}.items():
        p, eps = configs.CONFIGS[name]()
        assert (p.K, p.N, p.L, p.NT) == (K, N, L, NT) and eps.shape == (L * NT,)
    p, eps = configs.c3_ensemble(n_delta=4, n_amp=5, NT=10)
    assert (p.K, p.G, p.N) == (20, 20, 3)
    p, eps = configs.c4_dense450(N=12, K=3, NT=4)
    assert np.allclose(p.tgt @ p.tgt.conj().T, np.eye(3))
    assert np.max(np.abs(np.linalg.eigvalsh(p.H0[0]))) == pytest.approx(1.0)


def test_shard_partition_covers_all_trajectories():
    p, _ = configs.random_problem(K=11, N=3, L=2, NT=5, G=4, seed=3)
    seen = []
    for r in range(3):
        s = p.shard(r, 3)
        assert s.K_global == 11
        for k in range(s.K):
            g = s.gen_of_traj[k]
            seen.append((s.psi0[k].tobytes(), s.H0[g].tobytes()))
    full = [(p.psi0[k].tobytes(), p.H0[p.gen_of_traj[k]].tobytes()) for k in range(p.K)]
    assert seen == full


def test_c_oracle_matches_python_oracle():
    from oracle import c_oracle as co
    from scipy.linalg import expm
    rng = np.random.default_rng(0)
    for n, sc in [(3, 1e-3), (5, 0.1), (4, 0.5), (6, 1.5), (7, 4.0), (9, 30.0)]:
        A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) * sc
        E = expm(A)
        assert np.max(np.abs(co.expm(A) - E)) < 1e-12 * np.max(np.abs(E))
    D = np.diag([0.0, 1.0, 0.5]).astype(complex)
    for fn in (gb.SM, gb.RE, gb.SS):
        p, eps = configs.random_problem(K=4, N=3, L=2, NT=14, seed=5, functional=fn, gb_kind=1, gb_D=D,
                                        lambda_b=0.4, ja_kind=1, lambda_a=0.3, shaped=True, hermitian=False,
                                        G=2)
        r = go.evaluate_gradient(go.from_problem(p), eps)
        c = co.evaluate_gradient(p, eps)
        assert np.max(np.abs(r["G"] - c["G"])) < 1e-12 * np.max(np.abs(r["G"]))
        assert abs(r["J"] - c["J"]) < 1e-12
        assert np.max(np.abs(r["tau"] - c["tau"])) < 1e-12


def test_julia_shim_matches_header():
    """julia/GrapeB200.jl cannot run here (no Julia): check statically that its struct
    mirrors `grape_b200_problem` field for field and that it only ccalls declared symbols."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "grape_b200.h")).read()
    jl = open(os.path.join(root, "julia", "GrapeB200.jl")).read()
    body = hdr[hdr.index("typedef struct grape_b200_problem {"):hdr.index("} grape_b200_problem;")]
    c_fields = re.findall(r"^\s*(?:const\s+)?(int32_t|double)\s*(\*?)\s*(\w+)\s*;", body, flags=re.M)
    sbody = jl[jl.index("struct Problem"):]
    sbody = sbody[:sbody.index("\nend")]
    j_fields = re.findall(r"^\s*(\w+)::([\w{}]+)\s*$", sbody, flags=re.M)
    assert [f[2] for f in c_fields] == [f[0] for f in j_fields]
    cmap = {("int32_t", ""): "Int32", ("double", ""): "Float64", ("int32_t", "*"): "Ptr{Int32}",
            ("double", "*"): "Ptr{Float64}"}
    for (ct, star, name), (jn, jt) in zip(c_fields, j_fields):
        assert cmap[(ct, star)] == jt, (name, ct + star, jt)
    from grape.jl_b200 import _lib
    assert [f[0] for f in _lib.ProblemDesc._fields_] == [f[0] for f in j_fields]
    for sym in re.findall(r"ccall\(\(:(\w+),", jl):
        assert sym in _lib.SYMBOLS, sym


def test_economised_polynomial_table(lib_built):
    """csrc/dense.cuh econ_table (host-only entry point): for every degree m the weights g[j] of the Taylor terms give
    a polynomial within 1e-17 (+ coefficient rounding) of exp(-i x) on [-theta_m, theta_m] (mpmath, 40 digits), theta_m grows with m and is well
    above the radius the Taylor series of the same degree serves -- the reference's Cheby propagator
    (docs/src/tutorial.md:308, 432) in the monomial basis."""
    import ctypes as C
    import math

    import mpmath as mp

    from grape.jl_b200 import _lib
    lib = _lib.load()
    mp.mp.dps = 40
    prev = 0.0
    for m in range(2, 21):
        th = C.c_double()
        g = (C.c_double * (m + 1))()
        assert lib.grape_b200_econ_table(m, C.byref(th), g) == 0
        theta = th.value
        assert theta >= prev and theta <= 1.0
        prev = theta
        if theta == 0.0:
            continue
        taylor_radius = (2e-17 * math.factorial(m)) ** (1.0 / m)
        assert theta >= min(1.0, 1.5 * taylor_radius), (m, theta, taylor_radius)
        worst = mp.mpf(0)
        for i in range(0, 41):
            x = mp.mpf(theta) * (2 * mp.mpf(i) / 40 - 1)
            pm = sum(mp.mpf(g[j]) * (-1j * x) ** j / mp.factorial(j) for j in range(m + 1))
            worst = max(worst, abs(pm - mp.exp(-1j * x)))
        # truncation bound 1e-17 + the rounding of the weights to double (g_0 = 1 - O(1e-17) rounds to 1; g_1 x: up to
        # 1.1e-16 theta): all an order of magnitude below the rounding of one matrix-vector product
        assert worst <= 2.02e-17 + 1.2e-16 * theta, (m, theta, float(worst))
        assert abs(g[0] - 1.0) <= 1e-15 and all(0.9 < g[j] <= 1.0 + 1e-15 for j in range(m + 1))
    assert lib.grape_b200_econ_table(1, C.byref(th), g) != 0
    th12 = C.c_double()
    g12 = (C.c_double * 13)()
    lib.grape_b200_econ_table(12, C.byref(th12), g12)
    assert th12.value >= 0.525   # C4 / C5: ||H_n dt|| <= 0.525 -> degree 12 = four 3-term stages (Taylor: 15..16)


def test_flops_model_follows_the_series_tables(lib_built, monkeypatch):
    """grape.jl_b200/peaks.py (the executed-flops model behind `roofline.achieved`) mirrors the kernels' plans: with the
    economised series C3 takes 5 orders per gradient step and C4 degree 12 per dense step; GRAPE_B200_ECON=0 gives the
    Taylor orders (7 resp. 15..16)."""
    from grape.jl_b200 import configs, peaks
    p3, e3 = configs.c3_ensemble(n_delta=8, n_amp=8)
    p4, e4 = configs.c4_dense450(N=64, K=16, NT=40)
    monkeypatch.delenv("GRAPE_B200_ECON", raising=False)
    econ3 = peaks.executed_flops_split(p3, e3, schedule=3)
    econ4 = peaks.dense_flops_per_unit(p4, e4, 1)[3]
    monkeypatch.setenv("GRAPE_B200_ECON", "0")
    tay3 = peaks.executed_flops_split(p3, e3, schedule=3)
    tay4 = peaks.dense_flops_per_unit(p4, e4, 1)[3]
    assert econ3["contraction"] < 0.85 * tay3["contraction"] and econ3["formation"] <= tay3["formation"]
    assert 10.0 <= econ4 <= 12.0 and 13.0 <= tay4 <= 16.0
    th = peaks.econ_theta()
    assert th[12] >= 0.525 and th[5] >= 0.0065 and all(th[m] <= th[m + 1] for m in range(2, 20))
