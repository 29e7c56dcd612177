#!/bin/bash
# GPU session helper (not a test), round 2 session 13 (1 GPU): two-warp gradient blocks for shards -- affected tests, sweep.
TAG=${1:-r2_s13}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_real_symmetric.py tests/test_gpu_parity_segmented.py tests/test_gpu_parity_small.py \
    tests/test_golden.py tests/test_gpu_multi.py tests/test_gpu_optimize.py tests/test_amplitude_slots.py \
    "tests/test_gpu_parity_full_size.py::test_c3_full_size_all_trajectories_vs_c_oracle" "tests/test_gpu_parity_full_size.py::test_c1_full_size_vs_oracle" \
    -q -m gpu --timeout 400 --maxfail=20 > ${OUT}_pytest.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest.txt
tail -6 ${OUT}_pytest.txt
timeout 300 python profiles/scripts/r2_c3_sweep5.py > ${OUT}_c3_sweep.txt 2>&1
cat ${OUT}_c3_sweep.txt
