#!/bin/bash
# GPU session helper (not a test), round 2 session 17 (1 GPU): economised polynomial in the real-symmetric small path --
# small-path test files, C3 / C1 A/B against the Taylor series and the segment-length sweep, default bench.
TAG=${1:-r2_s17}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_real_symmetric.py tests/test_gpu_parity_segmented.py tests/test_gpu_parity_small.py tests/test_gpu_optimize.py tests/test_gpu_parity_full_size.py -q -m gpu --maxfail=10 --timeout 600 --durations=5 > ${OUT}_pytest_small.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_small.txt
tail -15 ${OUT}_pytest_small.txt
timeout 600 python profiles/scripts/r2_c3_sweep8.py > ${OUT}_c3_econ.txt 2>&1
cat ${OUT}_c3_econ.txt | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 5 --no-extra > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
tail -3 ${OUT}_bench.err
python - <<P
import json
d = json.loads(open("${OUT}_bench_c3.json").read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
P
