#!/bin/bash
# GPU session helper (not a test), round 2 session 23 (1 GPU): scan schedule of the sub-warp path (C2).
TAG=${1:-r2_s23}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_segmented.py tests/test_gpu_parity_warp.py tests/test_gpu_parity_full_size.py -q -m gpu --maxfail=10 --timeout 600 -k "warp or c2" > ${OUT}_pytest_warp.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_warp.txt
tail -25 ${OUT}_pytest_warp.txt
for scan in 1 0; do
  GRAPE_B200_WSEG_SCAN=$scan timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c2_scan${scan}.json 2>> ${OUT}_bench.err
done
for S in 8 12 24 32; do
  GRAPE_B200_SEG_S=$S timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c2_S${S}.json 2>> ${OUT}_bench.err
done
python - <<P
import json, glob
for f in sorted(glob.glob("${OUT}_bench_c2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"].get("phase_ms"))
    except Exception as e:
        print(f, "no result", e)
P
tail -3 ${OUT}_bench.err
