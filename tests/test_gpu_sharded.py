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


@pytest.mark.parametrize("fn", [gb.SS, gb.RE])
def test_single_exchange_for_uncoupled_functionals(lib_built, fn):
    """J_T_ss / J_T_re: chi_k depends on tau_k only, so the sums may be reduced AFTER the backward sweep, together with
    the gradient (DevicePipeline.step with more than one rank). Emulated here with two shards on one device: local
    forward + backward on each, one 'all-reduce' of {sums, grad_J_Tb}, enqueue_combine -> G and J of the full problem."""
    import torch
    from grape.jl_b200.engine import GrapeEngine
    from grape.jl_b200.sharded import DevicePipeline, _DevArray
    p, eps = configs.c3_ensemble(n_delta=4, n_amp=6, NT=70, functional=fn, ja_kind=1, lambda_a=0.05)
    e = GrapeEngine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    e.close()
    dev = torch.device("cuda", 0)
    d_eps = torch.from_numpy(eps).to(dev)
    shards = [GrapeEngine(p.shard(r, 2)) for r in range(2)]
    pipes = [DevicePipeline(s) for s in shards]
    assert all(not q.coupled for q in pipes)
    for s in shards:
        s.enqueue_forward(d_eps.data_ptr())
        s.enqueue_backward()
        s.finish()
    sums = sum(q.sums_t.clone() for q in pipes)
    gTb = sum(q.gTb_t.clone() for q in pipes)
    for s, q in zip(shards, pipes):
        q.sums_t.copy_(sums)
        q.gTb_t.copy_(gTb)
        torch.cuda.synchronize()
        s.enqueue_combine()
        s.finish()
        Jp = torch.as_tensor(_DevArray(s.device_ptr(4), 3), device=dev).cpu().numpy()
        assert np.max(np.abs(q.gradient().cpu().numpy() - G)) <= 1e-13 * np.max(np.abs(G))
        assert abs(float(np.sum(Jp)) - J) <= 1e-13
    for s in shards:
        s.close()
