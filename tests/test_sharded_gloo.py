"""world_size-2 gloo test of the trajectory-sharding host logic:
forward -> all-reduce(sums) -> backward -> all-reduce(gradient), with the
oracle standing in for the per-rank engine (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, functional, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import grape.jl_b200 as gb
    from grape.jl_b200 import configs
    from grape.jl_b200.sharded import ShardedGrape
    from tests.oracle_engine import OracleEngine
    D = np.diag([0.0, 1.0, 0.5])
    p, eps = configs.random_problem(K=7, N=3, L=2, NT=9, G=3, seed=17, functional=functional,
                                    gb_kind=gb.GB_QUADFORM, gb_D=D, lambda_b=0.3,
                                    ja_kind=gb.JA_FLUENCE, lambda_a=0.2,
                                    weights=np.linspace(0.5, 1.5, 7))
    sh = ShardedGrape(p, OracleEngine)
    G = np.zeros_like(eps)
    J = sh.evaluate_gradient(G, eps)
    Jf = sh.evaluate_functional(eps)
    np.save(os.path.join(out, f"G_{functional}_{rank}.npy"), np.concatenate([[J, Jf], G]))
    dist.destroy_process_group()


@pytest.mark.parametrize("functional", [0, 1, 2])
def test_two_rank_sharding_matches_single(tmp_path, functional):
    import grape.jl_b200 as gb
    from grape.jl_b200 import configs
    from oracle import grape_oracle as go
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), functional, str(tmp_path)), nprocs=world, join=True)
    D = np.diag([0.0, 1.0, 0.5])
    p, eps = configs.random_problem(K=7, N=3, L=2, NT=9, G=3, seed=17, functional=functional,
                                    gb_kind=gb.GB_QUADFORM, gb_D=D, lambda_b=0.3,
                                    ja_kind=gb.JA_FLUENCE, lambda_a=0.2,
                                    weights=np.linspace(0.5, 1.5, 7))
    ref = go.evaluate_gradient(go.from_problem(p), eps)
    outs = [np.load(tmp_path / f"G_{functional}_{r}.npy") for r in range(world)]
    assert np.array_equal(outs[0], outs[1])          # every rank holds the same J and gradient
    assert abs(outs[0][0] - ref["J"]) < 1e-12 and abs(outs[0][1] - ref["J"]) < 1e-12
    assert np.max(np.abs(outs[0][2:] - ref["G"])) < 1e-12 * np.max(np.abs(ref["G"]))
