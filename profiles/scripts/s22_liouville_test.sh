#!/bin/bash
# GPU session helper (not a test): Liouville-space trajectories through the small path
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity_small.py -q -m gpu -x -k "liouville" > ${OUT}_pytest_gpu_liouville.txt 2>&1; echo "pytest exit $?" >> ${OUT}_pytest_gpu_liouville.txt; tail -12 ${OUT}_pytest_gpu_liouville.txt
