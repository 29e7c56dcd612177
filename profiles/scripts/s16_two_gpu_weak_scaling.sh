#!/bin/bash
# GPU session helper (not a test): 2-GPU weak-scaling records (one rank per GPU, NCCL)
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
nvidia-smi -L > ${OUT}_gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 > ${OUT}_bench_c3_n2_weak.json 2> ${OUT}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 10 --impl reference > ${OUT}_bench_c3_n2_reference.json 2>> ${OUT}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --steps 4 --warmup 3 > ${OUT}_bench_c5_n2_weak.json 2>> ${OUT}_bench_n2.err
timeout 200 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c3_n1.json 2>> ${OUT}_bench_n2.err
for f in c3_n2_weak c5_n2_weak c3_n1; do python - <<P
import json
try:
    d=json.loads([l for l in open("${OUT}_bench_${f}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("${f}", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["workload"])
except Exception as e: print("${f}", "no result", e)
P
done
tail -5 ${OUT}_bench_n2.err
