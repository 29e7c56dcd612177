#!/bin/bash
# GPU session helper (not a test), round 2 session 26 (1 GPU): memcheck over the late round-2 kernels, final default
# bench (+ C4 / C5 extras), C1 / C2 benches, smoke().
TAG=${1:-r2_s26}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/scripts/_sanitize_r2b.py > ${OUT}_memcheck.txt 2>&1
echo "memcheck exit $?" >> ${OUT}_memcheck.txt
grep -E "ERROR SUMMARY|SANITIZE_R2B_DONE|memcheck exit|Invalid|Error" ${OUT}_memcheck.txt | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${OUT}_smoke.txt 2>&1; echo "smoke exit $?" >> ${OUT}_smoke.txt; tail -3 ${OUT}_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
echo "bench exit $?"; tail -3 ${OUT}_bench.err
timeout 200 python bench.py --workload c1 --steps 200 --warmup 10 > ${OUT}_bench_c1.json 2>> ${OUT}_bench.err
timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 > ${OUT}_bench_c2.json 2>> ${OUT}_bench.err
python - <<P
import json
for f in ("${OUT}_bench_c3.json", "${OUT}_bench_c1.json", "${OUT}_bench_c2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r.get("frac"), r.get("step_frac"), d.get("cpu_baseline", {}).get("value"))
        for k, v in (d.get("extra_workloads") or {}).items():
            print("  extra", k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), (v.get("roofline") or {}).get("step_frac"), v.get("error"))
    except Exception as e:
        print(f, "no result", e)
P
