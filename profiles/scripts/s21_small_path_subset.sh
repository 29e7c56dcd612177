#!/bin/bash
# GPU session helper (not a test): remaining small-path / optimizer / sharding tests on the final tree
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_optimize.py tests/test_gpu_sharded.py tests/test_gpu_parity_warp.py -q -m gpu -x > ${OUT}_pytest_gpu_subset.txt 2>&1; echo "pytest exit $?" >> ${OUT}_pytest_gpu_subset.txt; tail -4 ${OUT}_pytest_gpu_subset.txt
