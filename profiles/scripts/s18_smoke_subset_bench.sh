#!/bin/bash
# GPU session helper (not a test): last sanity check of the tree -- smoke(), the tests that touch the tau reduction at
# full C3 size and the sharded pipeline, default bench
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke > ${OUT}_smoke.txt 2>&1; echo "smoke exit $?" >> ${OUT}_smoke.txt; tail -7 ${OUT}_smoke.txt
timeout 200 python -m pytest tests -q -m gpu -x -k "c3_full_size or sharded or golden or optimize" > ${OUT}_pytest_gpu_subset.txt 2>&1; echo "pytest exit $?" >> ${OUT}_pytest_gpu_subset.txt; tail -4 ${OUT}_pytest_gpu_subset.txt
timeout 100 python bench.py > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
python - <<P
import json
d=json.loads(open("${OUT}_bench_c3.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_fp64"]["frac"], d["roofline"]["phase_ms"], d["gpu_launches"], d["clocks"])
P
