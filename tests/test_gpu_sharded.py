"""Device-resident sharded pipeline (grape.jl_b200/sharded.py) on one GPU: same numbers as the
blocking C-ABI call; two shards on one device reproduce the unsharded gradient."""
import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs

pytestmark = pytest.mark.gpu


def test_pipeline_matches_blocking_call(lib_built):
    import torch
    from grape.jl_b200.engine import GrapeEngine
    from grape.jl_b200.sharded import ShardedGrape
    for fn, kw in ((gb.SS, {}), (gb.SM, dict(ja_kind=1, lambda_a=0.05))):
        p, eps = configs.c3_ensemble(n_delta=6, n_amp=5, NT=90, functional=fn, **kw)
        e = GrapeEngine(p)
        G0 = np.zeros_like(eps)
        J0 = e.evaluate_gradient(G0, eps)
        sh = ShardedGrape(p, lambda lp: e, rank=0, world=1, device=torch.device("cuda", 0))
        assert sh._pipe is not None
        G1 = np.zeros_like(eps)
        J1 = sh.evaluate_gradient(G1, eps)
        assert np.array_equal(G0, G1) and J0 == J1
        sh.close()
        e.close()


def test_two_shards_on_one_device_sum_to_full_gradient(lib_built):
    from grape.jl_b200.engine import GrapeEngine
    p, eps = configs.c3_ensemble(n_delta=4, n_amp=6, NT=70, functional=gb.SM)
    e = GrapeEngine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    shards = [GrapeEngine(p.shard(r, 2)) for r in range(2)]
    sums = sum(s.forward(eps) for s in shards)
    Gs = np.zeros_like(eps)
    for s in shards:
        Gp = np.zeros_like(eps)
        Jp = s.backward(sums, Gp)
        Gs += Gp
    assert np.max(np.abs(Gs - G)) <= 1e-13 * np.max(np.abs(G))
    assert abs(np.sum(Jp) - J) <= 1e-13
