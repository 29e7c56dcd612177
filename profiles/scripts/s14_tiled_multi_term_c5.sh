#!/bin/bash
# GPU session helper (not a test): tiled multi-term chain -- targeted tests and C5 bench with 3 terms / 1 term per barrier
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_parity_dense_krylov.py -x -q -m gpu -k "several_terms or tiled_kernels or running_costs or c5_full_width or large_n" > ${OUT}_pytest_gpu_dense_tiled.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu_dense_tiled.txt
tail -15 ${OUT}_pytest_gpu_dense_tiled.txt
timeout 240 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > ${OUT}_bench_c5.json 2> ${OUT}_bench_c5.err
GRAPE_B200_DENSE_TERMS=1 timeout 240 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > ${OUT}_bench_c5_single_term.json 2>> ${OUT}_bench_c5.err
for f in c5 c5_single_term; do python - <<P
import json
try:
    d=json.loads(open("${OUT}_bench_${f}.json").read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("${f}", d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("gradient_form"), r.get("phase_ms"), r.get("step_frac"))
except Exception as e: print("${f}", "no result", e)
P
done
tail -5 ${OUT}_bench_c5.err
