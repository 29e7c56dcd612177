#!/bin/bash
# GPU session helper (not a test): segmented / sub-warp / optimizer test files on the final tree
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_gpu_parity_segmented.py tests/test_gpu_parity_warp.py tests/test_gpu_optimize.py -q -m gpu -x > ${OUT}_pytest_gpu_subset.txt 2>&1; echo "pytest exit $?" >> ${OUT}_pytest_gpu_subset.txt; tail -4 ${OUT}_pytest_gpu_subset.txt
