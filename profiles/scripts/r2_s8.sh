#!/bin/bash
# GPU session helper (not a test), round 2 session 8 (1 GPU): ncu --set full (+ source page) of the concurrent dense chain on
# C4 width; sanitizer memcheck over the round-2 kernels; small-path micro-optimisations (templated cos/sin degree,
# interleaved control reductions, scan schedule for small ensembles): affected tests, C1/C2/C3 benches, shard sweep.
TAG=${1:-r2_s8}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_real_symmetric.py tests/test_gpu_parity_segmented.py tests/test_gpu_parity_small.py \
    tests/test_golden.py tests/test_gpu_optimize.py tests/test_amplitude_slots.py tests/test_gpu_multi.py tests/test_gpu_sharded.py \
    "tests/test_gpu_parity_full_size.py::test_c1_full_size_vs_oracle" "tests/test_gpu_parity_full_size.py::test_c3_full_size_all_trajectories_vs_c_oracle" \
    -q -m gpu --timeout 400 --maxfail=20 > ${OUT}_pytest.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest.txt
tail -8 ${OUT}_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > ${OUT}_bench_c3.json 2> ${OUT}_bench.err
timeout 200 python bench.py --workload c1 --steps 200 --warmup 10 > ${OUT}_bench_c1.json 2>> ${OUT}_bench.err
timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 > ${OUT}_bench_c2.json 2>> ${OUT}_bench.err
timeout 300 python profiles/scripts/r2_c3_sweep3.py > ${OUT}_c3_sweep.txt 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/scripts/_sanitize_r2.py > ${OUT}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> ${OUT}_sanitizer_memcheck.txt
tail -4 ${OUT}_sanitizer_memcheck.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dense_chain" -c 1 \
    -f -o ${OUT}_ncu_c4 python profiles/scripts/_prof_dense.py c4 30 > ${OUT}_ncu_c4.log 2>&1
ncu -i ${OUT}_ncu_c4.ncu-rep --page raw --csv > ${OUT}_ncu_full_c4_chain_concurrent_raw.csv 2>/dev/null
ncu -i ${OUT}_ncu_c4.ncu-rep --page source --csv > ${OUT}_ncu_c4_source.csv 2>/dev/null
rm -f ${OUT}_ncu_c4.ncu-rep
python - <<P
import json
for f in ("${OUT}_bench_c3.json", "${OUT}_bench_c1.json", "${OUT}_bench_c2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r.get("frac"), r.get("step_frac"), r.get("phase_ms"))
    except Exception as e:
        print(f, "no result", e)
for l in open("${OUT}_c3_sweep.txt"):
    try:
        d = json.loads(l); print(d["K"], d["mode"], d.get("S"), round(d["ms"], 4), d["phases"][:5])
    except Exception:
        print(l.strip()[:300])
P
