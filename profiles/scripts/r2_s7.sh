#!/bin/bash
# GPU session helper (not a test), round 2 session 7 (1 GPU): split-phase barriers in the concurrent dense chains --
# dense test files, C4 bench (split on), compute-sanitizer memcheck over the round-2 kernels, full suite, default bench.
TAG=${1:-r2_s7}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_dense_krylov.py tests/test_gpu_parity_dense.py tests/test_gpu_parity_full_size.py -q -m gpu --timeout 400 --maxfail=10 > ${OUT}_pytest_dense.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_dense.txt
tail -12 ${OUT}_pytest_dense.txt
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > ${OUT}_bench_c4.json 2> ${OUT}_bench.err
echo "c4 exit $?"; tail -2 ${OUT}_bench.err
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/scripts/_sanitize_r2.py > ${OUT}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> ${OUT}_sanitizer_memcheck.txt
tail -6 ${OUT}_sanitizer_memcheck.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=25 --timeout 400 --deselect tests/test_gpu_parity_dense_krylov.py --deselect tests/test_gpu_parity_dense.py --deselect tests/test_gpu_parity_full_size.py > ${OUT}_pytest_rest.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_rest.txt
tail -6 ${OUT}_pytest_rest.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${OUT}_bench_c3.json 2>> ${OUT}_bench.err
python - <<P
import json
for f in ("${OUT}_bench_c4.json", "${OUT}_bench_c3.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
        for k, v in (d.get("extra_workloads") or {}).items():
            print("  extra", k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("step_frac"), v.get("error"))
    except Exception as e:
        print(f, "no result", e)
P
