#!/bin/bash
# GPU session helper (not a test): final record of a round -- full GPU suite, benches of all five configs (bench.py
# default = c3), reference arm, launch list of the default bench, ncu --set full of the C5 chain / contraction kernels
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > ${OUT}_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > ${OUT}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu.txt
tail -6 ${OUT}_pytest_gpu.txt
timeout 200 python bench.py > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > ${OUT}_bench_c3_reference.json 2>> ${OUT}_bench.err
timeout 100 python bench.py --workload c1 > ${OUT}_bench_c1.json 2>> ${OUT}_bench.err
timeout 100 python bench.py --workload c2 > ${OUT}_bench_c2.json 2>> ${OUT}_bench.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 > ${OUT}_bench_c4.json 2>> ${OUT}_bench.err
timeout 240 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > ${OUT}_bench_c5.json 2>> ${OUT}_bench.err
for f in c3 c1 c2 c4 c5; do python - <<P
import json
try:
    d=json.loads(open("${OUT}_bench_${f}.json").read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("${f}", d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), d.get("roofline_fp64",{}).get("frac"), r.get("gradient_form"), r.get("phase_ms"), r.get("step_frac"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("${f}", "no result", e)
P
done
tail -5 ${OUT}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${OUT}_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dense2_chain_multi|kry_contract" -c 3 \
    -f -o ${OUT}_ncu_c5 python tests/_prof_dense.py c5 6 > ${OUT}_ncu_c5.log 2>&1
ncu -i ${OUT}_ncu_c5.ncu-rep --page raw --csv > ${OUT}_ncu_full_c5_raw.csv 2>/dev/null
rm -f ${OUT}_ncu_c5.ncu-rep
ls -la gpurun_out | tail -15
