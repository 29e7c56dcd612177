"""torchrun worker of tests/test_gpu_multi.py::test_one_process_per_gpu_matches_single_gpu (one rank per GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import grape.jl_b200 as gb
    from grape.jl_b200 import configs
    from grape.jl_b200.engine import GrapeEngine
    from grape.jl_b200.sharded import ShardedGrape, DevicePipeline
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    exchange = os.environ.get("GRAPE_TEST_EXCHANGE", "p2p")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for functional in (gb.SS, gb.SM):
        p, eps = configs.c3_ensemble(n_delta=8, n_amp=16, NT=300, functional=functional, ja_kind=1, lambda_a=0.02)
        full = GrapeEngine(p, device=local)                       # the unsharded problem on this rank's own GPU
        G0 = np.zeros_like(eps)
        J0 = full.evaluate_gradient(G0, eps)
        full.close()
        sh = ShardedGrape(p, lambda lp: GrapeEngine(lp, device=local), rank=rank, world=world, device=dev,
                          exchange=exchange)
        assert sh._pipe.exchange == exchange, (sh._pipe.exchange, sh._pipe.exchange_note)
        G = np.zeros_like(eps)
        J = sh.evaluate_gradient(G, eps)                          # host API over ranks
        sc = np.max(np.abs(G0))
        assert abs(J - J0) <= 1e-12 and np.max(np.abs(G - G0)) <= 1e-12 * sc, (J - J0, np.max(np.abs(G - G0)) / sc)
        # device-resident pipeline, several back-to-back steps without host synchronisation
        stream = torch.cuda.ExternalStream(sh.engine.stream(), device=dev)
        d_eps = torch.from_numpy(eps).to(dev)
        with torch.cuda.stream(stream):
            for _ in range(5):
                sh._pipe.step(d_eps)
        sh._pipe.finish()
        Gd = sh._pipe.gradient().cpu().numpy()
        assert np.max(np.abs(Gd - G0)) <= 1e-12 * sc
        if exchange == "p2p":
            assert np.array_equal(Gd, G), "device pipeline and host API disagree"
        # identical on every rank
        t = torch.from_numpy(Gd.copy()).to(dev)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "gradient differs between ranks"
        if exchange == "p2p":
            Jf = sh.evaluate_functional(eps)
            assert abs(Jf - J0) <= 1e-12
        sh.close()
        sh.engine.close()
    dist.barrier()
    print("P2P_WORKER_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
