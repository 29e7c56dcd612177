#!/bin/bash
# GPU session helper (not a test), round 2 session 3 (1 GPU): full GPU suite after the scan prologue / one-pass chains /
# amplitude mode, default bench, schedule-rule sweep, launch list.
TAG=${1:-r2_s3}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=25 --timeout 400 --durations=5 > ${OUT}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_gpu.txt
tail -25 ${OUT}_pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
echo "bench exit $?"; tail -3 ${OUT}_bench.err
timeout 300 python profiles/scripts/r2_c3_sweep3.py > ${OUT}_c3_sweep.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${OUT}_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-sustained > /dev/null 2>&1
python - <<P
import json
for f in ("${OUT}_bench_c3.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
        for k, v in (d.get("extra_workloads") or {}).items():
            print("  extra", k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("step_frac"), v.get("error"))
    except Exception as e:
        print(f, "no result", e)
for l in open("${OUT}_c3_sweep.txt"):
    try:
        d = json.loads(l); print(d["K"], d["mode"], d.get("S"), round(d["ms"], 4), d["phases"][:5])
    except Exception:
        print(l.strip()[:300])
P
