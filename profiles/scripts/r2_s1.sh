#!/bin/bash
# GPU session helper (not a test), round 2 session 1 (1 GPU): full GPU suite incl. the new full-size parity, multi-shard
# and saddle tests; default bench line (+ C4/C5 extras) and reference arm; C3 shard-size / segment-length sweep; launch
# list + ncu --set full of the staged real-symmetric kernels.
TAG=${1:-r2_s1}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > ${OUT}_gpu.txt 2>&1
nproc >> ${OUT}_gpu.txt
timeout 1200 python -m pytest tests -q -m gpu --maxfail=20 --timeout 400 --durations=15 > ${OUT}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu.txt
tail -30 ${OUT}_pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
echo "bench exit $?"; tail -3 ${OUT}_bench.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > ${OUT}_bench_c3_reference.json 2>> ${OUT}_bench.err
timeout 300 python profiles/scripts/r2_c3_sweep.py > ${OUT}_c3_sweep.txt 2>&1
tail -5 ${OUT}_c3_sweep.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${OUT}_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"small_seggrad_sym2|small_formseg_sym2" -c 2 \
    -f -o ${OUT}_ncu_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > ${OUT}_ncu_c3.log 2>&1
ncu -i ${OUT}_ncu_c3.ncu-rep --page raw --csv > ${OUT}_ncu_full_c3_sym2_raw.csv 2>/dev/null
ncu -i ${OUT}_ncu_c3.ncu-rep --page source --csv > ${OUT}_ncu_c3_sym2_source.csv 2>/dev/null
ls -la gpurun_out | tail -12
python - <<P
import json
for f in ("${OUT}_bench_c3.json", "${OUT}_bench_c3_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
        for k, v in (d.get("extra_workloads") or {}).items():
            print("  extra", k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("step_frac"), v.get("error"))
    except Exception as e:
        print(f, "no result", e)
P
