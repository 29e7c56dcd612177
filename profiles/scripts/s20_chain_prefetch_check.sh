#!/bin/bash
# GPU session helper (not a test): boundary chains with chunked propagator loads -- segmented / real-symmetric tests + default bench
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity_segmented.py tests/test_gpu_parity_real_symmetric.py tests/test_golden.py -q -m gpu -x -k "not warp" > ${OUT}_pytest_gpu_subset.txt 2>&1; echo "pytest exit $?" >> ${OUT}_pytest_gpu_subset.txt; tail -4 ${OUT}_pytest_gpu_subset.txt
timeout 60 python bench.py --no-cpu-baseline > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
python - <<P
import json
d=json.loads(open("${OUT}_bench_c3.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_fp64"]["frac"], d["roofline"]["phase_ms"], d["gpu_launches"])
P
