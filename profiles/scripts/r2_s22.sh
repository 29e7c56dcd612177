#!/bin/bash
# GPU session helper (not a test), round 2 session 22 (1 GPU): final single-GPU record with the economised kernels -- full GPU suite, smoke(), default
# bench (+ C4/C5 extras), reference arm, C1/C2 benches, launch list, ncu --set full of the boundary chain.
TAG=${1:-r2_s22}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=25 --timeout 400 --durations=5 > ${OUT}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu.txt
tail -12 ${OUT}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${OUT}_smoke.txt 2>&1; echo "smoke exit $?" >> ${OUT}_smoke.txt; tail -8 ${OUT}_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
echo "bench exit $?"; tail -3 ${OUT}_bench.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > ${OUT}_bench_c3_reference.json 2>> ${OUT}_bench.err
timeout 200 python bench.py --workload c1 --steps 200 --warmup 10 > ${OUT}_bench_c1.json 2>> ${OUT}_bench.err
timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 > ${OUT}_bench_c2.json 2>> ${OUT}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${OUT}_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"small_segchain_dual|small_seggrad_sym2|small_formseg_sym2" -c 3 \
    -f -o ${OUT}_ncu_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > ${OUT}_ncu_c3.log 2>&1
ncu -i ${OUT}_ncu_c3.ncu-rep --page raw --csv > ${OUT}_ncu_full_c3_raw.csv 2>/dev/null
ncu -i ${OUT}_ncu_c3.ncu-rep --page source --csv > ${OUT}_ncu_c3_source.csv 2>/dev/null
rm -f ${OUT}_ncu_c3.ncu-rep
python - <<P
import json
for f in ("${OUT}_bench_c3.json", "${OUT}_bench_c3_reference.json", "${OUT}_bench_c1.json", "${OUT}_bench_c2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
        for k, v in (d.get("extra_workloads") or {}).items():
            print("  extra", k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("step_frac"), v.get("error"))
    except Exception as e:
        print(f, "no result", e)
P
