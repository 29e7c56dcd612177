#!/bin/bash
# GPU session helper (not a test), round 2 session 10 (1 GPU): boundary chain with one warp per direction -- affected tests,
# default bench, launch list.
TAG=${1:-r2_s10}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_real_symmetric.py tests/test_gpu_parity_segmented.py tests/test_gpu_parity_small.py \
    tests/test_golden.py tests/test_gpu_multi.py "tests/test_gpu_parity_full_size.py::test_c3_full_size_all_trajectories_vs_c_oracle" \
    -q -m gpu --timeout 400 --maxfail=20 > ${OUT}_pytest.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest.txt
tail -6 ${OUT}_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${OUT}_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > /dev/null 2>&1
python - <<P
import json
d = json.loads(open("${OUT}_bench_c3.json").read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
P
grep -c . ${OUT}_launches_c3.csv
