"""Multi-GPU behind the C-ABI (include/grape_b200.h: grape_b200_multi_*, grape_b200_xchg_*; csrc/xchg.cuh): the
shards of a trajectory-sharded problem reduce the tau sums and the gradient themselves over peer memory.

  * one process, several shards -- `grape_b200_multi_create`.  With devices = [0, 0, ...] the shards share ONE
    GPU (their exchange kernels run concurrently on different streams), so the exchange protocol is covered by
    the single-GPU test tier; with >= 2 GPUs the same tests run on distinct devices over NVLink.
  * one process per GPU -- CUDA IPC handles passed around once, then no collective library in the step
    (tests/_p2p_worker.py under torchrun; needs >= 2 GPUs).

SURVEY gate G6: results within 1e-12 of the 1-GPU result, bit-identical run to run and on every rank."""
import os
import subprocess
import sys

import numpy as np
import pytest

import grape.jl_b200 as gb
from grape.jl_b200 import configs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _single(p, eps):
    from grape.jl_b200.engine import GrapeEngine
    e = GrapeEngine(p)
    G = np.zeros_like(eps)
    J = e.evaluate_gradient(G, eps)
    out = dict(J=J, G=G, tau=e.tau_vals.copy(), J_parts=e.J_parts.copy(), Jf=e.evaluate_functional(eps),
               fs=e.final_states())
    e.close()
    return out


def _check_multi(p, eps, devices):
    from grape.jl_b200.engine import MultiGrapeEngine
    ref = _single(p, eps)
    m = MultiGrapeEngine(p, devices)
    assert m.size() == len(devices)
    G = np.zeros_like(eps)
    J = m.evaluate_gradient(G, eps)
    sc = max(np.max(np.abs(ref["G"])), 1e-300)
    assert abs(J - ref["J"]) <= 1e-12
    assert np.max(np.abs(m.J_parts - ref["J_parts"])) <= 1e-12
    assert np.max(np.abs(G - ref["G"])) <= 1e-12 * sc
    # per-trajectory work is unchanged by the split up to the schedule the shard size selects (segment length, scan)
    assert np.max(np.abs(m.tau_vals - ref["tau"])) <= 1e-13
    assert abs(m.evaluate_functional(eps) - ref["Jf"]) <= 1e-12
    assert np.max(np.abs(m.final_states() - ref["fs"])) <= 1e-13
    for _ in range(3):                                             # fixed-order reductions: bit-identical run to run
        G2 = np.zeros_like(eps)
        J2 = m.evaluate_gradient(G2, eps)
        assert J2 == J and np.array_equal(G2, G)
    x = eps * 1.01                                                 # and the epochs keep advancing correctly
    ref2 = _single(p, x)
    J3 = m.evaluate_gradient(G, x)
    assert abs(J3 - ref2["J"]) <= 1e-12 and np.max(np.abs(G - ref2["G"])) <= 1e-12 * sc
    m.close()


@pytest.mark.parametrize("functional", [gb.SS, gb.SM, gb.RE])
@pytest.mark.parametrize("nshard", [2, 3])
def test_shards_on_one_device_small_path(lib_built, functional, nshard):
    p, eps = configs.c3_ensemble(n_delta=5, n_amp=7, NT=90, functional=functional, ja_kind=1, lambda_a=0.05)
    _check_multi(p, eps, [0] * nshard)


def test_shards_on_one_device_shared_generators_weights_costs(lib_built):
    D = np.diag([0.0, 1.0, 0.5]).astype(complex)
    p, eps = configs.random_problem(K=11, N=3, L=2, NT=14, G=4, seed=5, functional=gb.SM, gb_kind=1, gb_D=D,
                                    lambda_b=0.4, ja_kind=1, lambda_a=0.3, shaped=True, hermitian=False,
                                    weights=np.linspace(0.5, 1.5, 11))
    _check_multi(p, eps, [0, 0, 0])
    p, eps = configs.random_problem(K=6, N=7, L=2, NT=12, G=1, seed=6, functional=gb.SS)   # sub-warp path
    p.tlist = p.tlist * 0.5
    _check_multi(p, eps, [0, 0])


def test_multi_more_devices_than_trajectories_is_an_error(lib_built):
    from grape.jl_b200.engine import MultiGrapeEngine, GrapeError
    p, eps = configs.c1_readme(NT=20)
    with pytest.raises(GrapeError, match="fewer trajectories than devices"):
        MultiGrapeEngine(p, [0, 0])


def test_multi_handle_on_distinct_devices(lib_built):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for functional in (gb.SS, gb.SM):
        p, eps = configs.c3_ensemble(n_delta=8, n_amp=8, NT=200, functional=functional)
        _check_multi(p, eps, list(range(min(n, 4))))
    p, eps = configs.c4_dense450(N=48, K=16, NT=6)                  # dense path: 8 columns per GPU
    _check_multi(p, eps, [0, 1])


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_one_process_per_gpu_matches_single_gpu(lib_built, exchange):
    """torchrun x 2: IPC-attached shards (or the NCCL path) against the unsharded 1-GPU evaluation, 1e-12, identical
    on both ranks and run to run."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, GRAPE_TEST_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617",
                        os.path.join(ROOT, "tests", "_p2p_worker.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("P2P_WORKER_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
