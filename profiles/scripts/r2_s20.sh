#!/bin/bash
# GPU session helper (not a test), round 2 session 20 (1 GPU): new defaults of the economised C3 kernels (128-register
# gradient kernel, two-step formation loop): small-path tests, A/B against GRAPE_B200_SYM_OCC=10, ncu --set full.
TAG=${1:-r2_s20}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_real_symmetric.py tests/test_gpu_parity_segmented.py tests/test_gpu_optimize.py -q -m gpu --maxfail=10 --timeout 600 > ${OUT}_pytest_small.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_small.txt
tail -4 ${OUT}_pytest_small.txt
timeout 600 python profiles/scripts/r2_c3_sweep9.py > ${OUT}_c3_variants.txt 2>&1
cat ${OUT}_c3_variants.txt | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"small_seggrad_sym2|small_formseg_sym2" -c 2 \
    -f -o ${OUT}_ncu_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > ${OUT}_ncu_c3.log 2>&1
ncu -i ${OUT}_ncu_c3.ncu-rep --page raw --csv > ${OUT}_ncu_full_c3_raw.csv 2>/dev/null
ncu -i ${OUT}_ncu_c3.ncu-rep --page source --csv > ${OUT}_ncu_c3_source.csv 2>/dev/null
rm -f ${OUT}_ncu_c3.ncu-rep
