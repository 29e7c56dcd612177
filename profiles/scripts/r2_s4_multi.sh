#!/bin/bash
# GPU session helper (not a test), N GPUs of one box: multi-GPU tests (multi handle on distinct devices, IPC-attached
# ranks under torchrun), then bench.py at N ranks: strong scaling of configs[2] (+ weak variant), in-library NVLink
# exchange vs the NCCL path.
N=${1:-2}
TAG=${2:-r2_s4}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > ${OUT}_gpus.txt 2>&1
nvidia-smi topo -m >> ${OUT}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu --timeout 600 > ${OUT}_pytest_multi.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_multi.txt
tail -15 ${OUT}_pytest_multi.txt
for n in $(seq 2 $N | awk -v N=$N '{ if ($1==2 || $1==4 || $1==8) print $1 }'); do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
      bench.py --gpus $n --steps 50 --warmup 5 > ${OUT}_bench_c3_n${n}.json 2> ${OUT}_bench_n${n}.err
  echo "bench n=$n exit $?"; tail -2 ${OUT}_bench_n${n}.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 --exchange nccl > ${OUT}_bench_c3_n${n}_nccl.json 2>> ${OUT}_bench_n${n}.err
done
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-extra --no-cpu-baseline > ${OUT}_bench_c3_n1.json 2> ${OUT}_bench_n1.err
python - <<P
import json, glob
for f in sorted(glob.glob("${OUT}_bench_c3_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        w = d.get("weak") or {}
        print(f, "N", d["n_gpus"], d["scaling"], "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"],
              "| weak value %.4g ms %.4f" % (w.get("value", 0), w.get("ms_per_step", 0)), d["config"].get("exchange"), d.get("parity_guard"))
    except Exception as e:
        print(f, "no result", e)
P
