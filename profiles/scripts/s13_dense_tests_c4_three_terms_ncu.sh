#!/bin/bash
# GPU session helper (not a test): dense Krylov tests, C4 bench with 3 / 2 / 2-in-kernel-formation Taylor terms per barrier, ncu of the chain, C5 bench
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_dense_krylov.py tests/test_gpu_parity_dense.py -q -m gpu > ${OUT}_pytest_gpu_dense.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu_dense.txt
tail -15 ${OUT}_pytest_gpu_dense.txt
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > ${OUT}_bench_c4.json 2> ${OUT}_bench_c4.err
GRAPE_B200_DENSE_TERMS=2 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > ${OUT}_bench_c4_two_terms.json 2>> ${OUT}_bench_c4.err
for f in c4 c4_two_terms; do python - <<P
import json
d=json.loads(open("${OUT}_bench_${f}.json").read().strip().splitlines()[-1])
r=d.get("roofline",{})
print("${f}", d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("gradient_form"), r.get("phase_ms"), r.get("step_frac"))
P
done
tail -3 ${OUT}_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_chain" -c 2 \
    -f -o ${OUT}_ncu_c4_chain python tests/_prof_dense.py c4 30 > ${OUT}_ncu_c4_chain.log 2>&1
ncu -i ${OUT}_ncu_c4_chain.ncu-rep --page raw --csv > ${OUT}_ncu_full_c4_chain_raw.csv 2>/dev/null
ncu -i ${OUT}_ncu_c4_chain.ncu-rep --page source --csv --kernel-name regex:dense_chain > ${OUT}_ncu_c4_chain_source.csv 2>/dev/null
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > ${OUT}_bench_c5.json 2> ${OUT}_bench_c5.err
python - <<P
import json
d=json.loads(open("${OUT}_bench_c5.json").read().strip().splitlines()[-1])
r=d.get("roofline",{})
print("c5", d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("gradient_form"), r.get("phase_ms"), r.get("step_frac"))
P
ls -la gpurun_out | tail -12
