#!/bin/bash
# GPU session helper (not a test), round 2 session 25 (1 GPU): radix-4 scan levels of the sub-warp path (C2).
TAG=${1:-r2_s25}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_segmented.py -q -m gpu --maxfail=5 --timeout 300 -k "warp_scan or warp_matches or warp_functional" > ${OUT}_pytest_warp_scan.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_warp_scan.txt
tail -6 ${OUT}_pytest_warp_scan.txt
timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c2.json 2>> ${OUT}_bench.err
GRAPE_B200_WSEG_RADIX=2 timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c2_radix2.json 2>> ${OUT}_bench.err
for S in 4 6 12; do
  GRAPE_B200_SEG_S=$S timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c2_S${S}.json 2>> ${OUT}_bench.err
done
python - <<P
import json, glob
for f in sorted(glob.glob("${OUT}_bench_c2*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"].get("phase_ms"))
    except Exception as e:
        print(f, "no result", e)
P
tail -3 ${OUT}_bench.err
