#!/bin/bash
# GPU session helper (not a test): dense-path tests, C3 segment-length sweep, C4 bench with the two-term strip chain on/off, ncu of the chain
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -k "dense or real_symmetric or segmented" > ${OUT}_pytest_gpu_subset.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu_subset.txt
tail -15 ${OUT}_pytest_gpu_subset.txt
python tests/_quick_c3.py 0 25 38 56 64 77 100 > ${OUT}_c3_segment_length_sweep.txt 2>&1; cat ${OUT}_c3_segment_length_sweep.txt
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > ${OUT}_bench_c4.json 2> ${OUT}_bench_c4.err
GRAPE_B200_DENSE_DUAL=0 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > ${OUT}_bench_c4_single_term.json 2>> ${OUT}_bench_c4.err
for f in c4 c4_single_term; do python - <<P
import json
d=json.loads(open("${OUT}_bench_${f}.json").read().strip().splitlines()[-1])
r=d.get("roofline",{})
print("${f}", d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("gradient_form"), r.get("phase_ms"), r.get("step_frac"))
P
done
tail -3 ${OUT}_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_chain" -c 2 \
    -f -o ${OUT}_ncu_c4_chain_dual python tests/_prof_dense.py c4 30 > ${OUT}_ncu_c4_chain_dual.log 2>&1
ncu -i ${OUT}_ncu_c4_chain_dual.ncu-rep --page raw --csv > ${OUT}_ncu_full_c4_chain_dual_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
