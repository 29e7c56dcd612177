#!/bin/bash
# GPU session helper (not a test), round 2 session 28 (2 GPUs): the in-library NVLink exchange with the final kernels.
TAG=${1:-r2_s28_2gpu}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 30 --warmup 5 > ${OUT}_bench_c3_n2.json 2> ${OUT}_bench_n2.err
echo "bench n=2 exit $?"; tail -2 ${OUT}_bench_n2.err
python - <<P
import json
d = json.loads(open("${OUT}_bench_c3_n2.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
print("N", d["n_gpus"], d["scaling"], "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "parity", d.get("parity_guard"))
print("weak", w.get("value"), w.get("ms_per_step"), (w.get("e2e") or {}).get("value"))
P
