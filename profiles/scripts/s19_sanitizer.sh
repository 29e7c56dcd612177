#!/bin/bash
# GPU session helper (not a test): compute-sanitizer memcheck + racecheck of the schedules added in sessions 11-14
TAG=${1:-sX}
OUT=gpurun_out/${TAG}
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/_sanitize_new.py > ${OUT}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> ${OUT}_sanitizer_memcheck.txt
tail -6 ${OUT}_sanitizer_memcheck.txt
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 3 python tests/_sanitize_new.py > ${OUT}_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" >> ${OUT}_sanitizer_racecheck.txt
tail -6 ${OUT}_sanitizer_racecheck.txt
