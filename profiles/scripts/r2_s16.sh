#!/bin/bash
# GPU session helper (not a test), round 2 session 16 (1 GPU): economised polynomial in the dense Krylov-form chains --
# dense test files, full-size C4 / C5 goldens, C4 / C5 bench with GRAPE_B200_ECON=1 (default) and 0.
TAG=${1:-r2_s16}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_dense_krylov.py tests/test_gpu_parity_full_size.py tests/test_gpu_parity_dense.py -q -m gpu --maxfail=10 --timeout 600 --durations=5 > ${OUT}_pytest_dense.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_dense.txt
tail -15 ${OUT}_pytest_dense.txt
for econ in 1 0; do
  for w in c4 c5; do
    GRAPE_B200_ECON=$econ timeout 400 python bench.py --workload $w --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-sustained > ${OUT}_bench_${w}_econ${econ}.json 2>> ${OUT}_bench.err
  done
done
python - <<P
import json
for econ in (1, 0):
    for w in ("c4", "c5"):
        f = "${OUT}_bench_%s_econ%d.json" % (w, econ)
        try:
            d = json.loads(open(f).read().strip().splitlines()[-1])
            r = d.get("roofline", {})
            print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
        except Exception as e:
            print(f, "no result", e)
P
tail -5 ${OUT}_bench.err
