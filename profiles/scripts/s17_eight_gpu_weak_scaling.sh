#!/bin/bash
# GPU session helper (not a test): 8-GPU / 4-GPU weak-scaling records (one rank per GPU, NCCL over NVLink)
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
nvidia-smi -L > ${OUT}_gpus.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 300 --warmup 10 > ${OUT}_bench_c3_n8_weak.json 2> ${OUT}_bench.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 300 --warmup 10 > ${OUT}_bench_c3_n4_weak.json 2>> ${OUT}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload c5 --steps 4 --warmup 3 > ${OUT}_bench_c5_n8_weak.json 2>> ${OUT}_bench.err
for f in c3_n8_weak c3_n4_weak c5_n8_weak; do python - <<P
import json
try:
    d=json.loads([l for l in open("${OUT}_bench_${f}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("${f}", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["workload"])
except Exception as e: print("${f}", "no result", e)
P
done
grep -v "^\*\|OMP_NUM\|^$" ${OUT}_bench.err | tail -5
