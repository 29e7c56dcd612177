#!/bin/bash
# GPU session helper (not a test): full GPU suite, C1-C3 benches with A/B switches of the small path,
# launch list and one ncu --set full capture of the real-symmetric kernels.  Usage: gpurun -- bash tests/_gpu_session.sh TAG
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > ${OUT}_gpu.txt 2>&1
timeout 1700 python -m pytest tests -q -m gpu > ${OUT}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu.txt
tail -5 ${OUT}_pytest_gpu.txt
python bench.py --workload c3 > ${OUT}_bench_c3.json 2> ${OUT}_bench_c3.err; tail -c 600 ${OUT}_bench_c3.json
GRAPE_B200_SYM_OCC=2 python bench.py --workload c3 --no-cpu-baseline > ${OUT}_bench_c3_occ2.json 2>> ${OUT}_bench_c3.err
GRAPE_B200_SEG_REAL=0 python bench.py --workload c3 --no-cpu-baseline > ${OUT}_bench_c3_hermitian.json 2>> ${OUT}_bench_c3.err
python bench.py --workload c1 > ${OUT}_bench_c1.json 2>> ${OUT}_bench_c3.err
python bench.py --workload c2 > ${OUT}_bench_c2.json 2>> ${OUT}_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > ${OUT}_bench_c3_reference.json 2>> ${OUT}_bench_c3.err
for f in c3 c3_occ2 c3_hermitian c1 c2; do python - <<P
import json
d=json.loads(open("${OUT}_bench_${f}.json").read().strip().splitlines()[-1])
print("${f}", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("roofline_fp64",{}).get("frac"), d.get("roofline",{}).get("phase_ms"))
P
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${OUT}_launches_c3.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"small_seggrad_sym|small_formseg_sym" -s 4 -c 2 \
    -f -o ${OUT}_ncu_c3_sym python tests/_quick_c3.py > ${OUT}_ncu_c3_sym.log 2>&1
ncu -i ${OUT}_ncu_c3_sym.ncu-rep --page raw --csv > ${OUT}_ncu_full_c3_sym_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
