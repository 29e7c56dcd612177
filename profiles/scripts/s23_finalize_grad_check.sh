#!/bin/bash
# GPU session helper (not a test): slab-parallel finalize_grad -- cross-section of the GPU tests + default bench
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_parity_real_symmetric.py tests/test_golden.py tests/test_gpu_sharded.py tests/test_gpu_parity_dense.py -q -m gpu -x > ${OUT}_pytest_gpu_subset.txt 2>&1; echo "pytest exit $?" >> ${OUT}_pytest_gpu_subset.txt; tail -4 ${OUT}_pytest_gpu_subset.txt
timeout 40 python bench.py --no-cpu-baseline > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
python - <<P
import json
d=json.loads(open("${OUT}_bench_c3.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"])
P
