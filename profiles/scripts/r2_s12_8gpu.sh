#!/bin/bash
# GPU session helper (not a test), one 8 x B200 box, final tree: multi-GPU tests, bench.py at 1 / 2 / 4 / 8 ranks
# (strong scaling of configs[2] as written + the weak variant; in-library NVLink exchange), reference arm.
TAG=${1:-r2_s12}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu --timeout 500 > ${OUT}_pytest_multi.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_multi.txt
tail -4 ${OUT}_pytest_multi.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra > ${OUT}_bench_c3_n1.json 2> ${OUT}_bench_n1.err
timeout 200 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > ${OUT}_bench_c3_n1_reference.json 2>> ${OUT}_bench_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
      bench.py --gpus $n --steps 20 --warmup 5 > ${OUT}_bench_c3_n${n}.json 2> ${OUT}_bench_n${n}.err
  echo "bench n=$n exit $?"; tail -2 ${OUT}_bench_n${n}.err
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
