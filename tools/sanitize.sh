#!/bin/bash
# compute-sanitizer over every kernel at reduced sizes.  ONE tool per gpurun call
# (B200_PROFILING.md: several tools in one call have wedged a GPU on this pool):
#   gpurun -- 'bash tools/sanitize.sh memcheck'      gpurun -- 'bash tools/sanitize.sh racecheck'
# Results: gpurun_out/sanitize_<tool>.txt (the summary goes into profiles/).
set -u
TOOL=${1:-memcheck}
OUT=gpurun_out/sanitize_${TOOL}.txt
mkdir -p gpurun_out
# the plain run must be green first (never run a sanitizer on a program that faults)
python tools/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1 && grep -q SANITIZE_RUN_OK gpurun_out/sanitize_plain.log || { echo "plain run failed" | tee $OUT; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 40 python tools/sanitize_run.py > $OUT 2>&1
echo "exit code $?" >> $OUT
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_RUN_OK|exit code|faces|spans" $OUT | tail -8
