#!/bin/bash
# GPU session helper (not a test), round 2 session 24 (1 GPU): sub-warp scan schedule with per-level kernels and the
# prefetching segment product: tests, C2 bench, segment-length sweep, per-kernel launch list.
TAG=${1:-r2_s24}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_segmented.py tests/test_gpu_parity_warp.py -q -m gpu --maxfail=10 --timeout 600 -k "warp and not full_size" > ${OUT}_pytest_warp.txt 2>&1
echo "pytest exit $?" >> ${OUT}_pytest_warp.txt
tail -8 ${OUT}_pytest_warp.txt
timeout 200 python bench.py --workload c2 --steps 200 --warmup 10 --no-cpu-baseline > ${OUT}_bench_c2.json 2>> ${OUT}_bench.err
for S in 8 12 24; do
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
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${OUT}_launches_c2.csv \
    python bench.py --workload c2 --steps 2 --warmup 2 --no-cpu-baseline --no-extra --no-sustained > /dev/null 2>&1
python - <<P
import csv, collections
rows = list(csv.reader(l for l in open("${OUT}_launches_c2.csv") if not l.startswith("==")))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi: continue
    agg.setdefault(r[ki][:60], []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1000:.2f} us")
P
tail -3 ${OUT}_bench.err
