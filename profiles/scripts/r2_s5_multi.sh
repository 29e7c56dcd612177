#!/bin/bash
# GPU session helper (not a test), N GPUs: the small-path changes of this session (ring chain, finalize fusion, exchange
# kernels without extra launches) -- affected test files, multi-GPU tests, bench at 1..N ranks (strong + weak, p2p / nccl).
N=${1:-2}
TAG=${2:-r2_s5}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_parity_real_symmetric.py \
    tests/test_gpu_parity_segmented.py tests/test_gpu_parity_small.py tests/test_gpu_parity_full_size.py tests/test_golden.py \
    tests/test_gpu_optimize.py tests/test_amplitude_slots.py tests/test_gate_functional.py tests/test_gpu_parity_warp.py \
    -q -m gpu --timeout 600 --maxfail=20 > ${OUT}_pytest.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest.txt
tail -15 ${OUT}_pytest.txt
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-extra > ${OUT}_bench_c3_n1.json 2> ${OUT}_bench_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
        bench.py --gpus $n --steps 50 --warmup 5 > ${OUT}_bench_c3_n${n}.json 2> ${OUT}_bench_n${n}.err
    echo "bench n=$n exit $?"; tail -2 ${OUT}_bench_n${n}.err
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --gpus $n --steps 50 --warmup 5 --exchange nccl > ${OUT}_bench_c3_n${n}_nccl.json 2>> ${OUT}_bench_n${n}.err
  fi
done
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
