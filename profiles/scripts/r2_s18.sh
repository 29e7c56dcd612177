#!/bin/bash
# GPU session helper (not a test), round 2 session 18 (1 GPU): register / inline-order variants of the economised
# gradient kernel, ncu --set full of the two heavy C3 kernels.
TAG=${1:-r2_s18}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 600 python profiles/scripts/r2_c3_sweep9.py > ${OUT}_c3_variants.txt 2>&1
cat ${OUT}_c3_variants.txt | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"small_seggrad_sym2|small_formseg_sym2" -c 2 \
    -f -o ${OUT}_ncu_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > ${OUT}_ncu_c3.log 2>&1
ncu -i ${OUT}_ncu_c3.ncu-rep --page raw --csv > ${OUT}_ncu_full_c3_raw.csv 2>/dev/null
ncu -i ${OUT}_ncu_c3.ncu-rep --page source --csv > ${OUT}_ncu_c3_source.csv 2>/dev/null
rm -f ${OUT}_ncu_c3.ncu-rep
ls -la gpurun_out | tail -5
