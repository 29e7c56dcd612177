#!/bin/bash
# GPU session helper (not a test), round 2 session 29 (8 GPUs): the in-library NVLink exchange with the final kernels.
TAG=${1:-r2_s29_8gpu}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
    bench.py --gpus 8 --steps 30 --warmup 5 > ${OUT}_bench_c3_n8.json 2> ${OUT}_bench_n8.err
echo "bench n=8 exit $?"; tail -2 ${OUT}_bench_n8.err
python - <<P
import json
d = json.loads(open("${OUT}_bench_c3_n8.json").read().strip().splitlines()[-1])
w = d.get("weak") or {}
print("N", d["n_gpus"], d["scaling"], "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "parity", d.get("parity_guard"))
print("weak", w.get("value"), w.get("ms_per_step"), (w.get("e2e") or {}).get("value"))
P
