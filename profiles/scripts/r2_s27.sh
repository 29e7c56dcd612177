#!/bin/bash
# GPU session helper (not a test), round 2 session 27 (1 GPU): full GPU suite on the final tree.
TAG=${1:-r2_s27}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=25 --timeout 400 --durations=8 -x > ${OUT}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu.txt
tail -16 ${OUT}_pytest_gpu.txt
