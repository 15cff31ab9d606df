#!/bin/bash
# compute-sanitizer over the kernel parity tests at reduced sizes.  One tool per gpurun call
# (B200_PROFILING.md: several tools in one call have wedged a GPU on this pool):
#   gpurun -- 'bash tools/sanitize.sh memcheck'      gpurun -- 'bash tools/sanitize.sh racecheck'
# Results: gpurun_out/sanitize_<tool>.txt (copy the summary into profiles/).
set -u
TOOL=${1:-memcheck}
OUT=gpurun_out/sanitize_${TOOL}.txt
mkdir -p gpurun_out
TESTS="tests/test_gpu_kernels.py::test_resize_bit_exact tests/test_gpu_kernels.py::test_letterbox_bit_exact tests/test_gpu_kernels.py::test_scrfd_heads_match_oracle tests/test_gpu_kernels.py::test_arcface_embeddings_match_oracle tests/test_gpu_kernels.py::test_decode_nms_bit_exact tests/test_gpu_kernels.py::test_align_chips_bit_exact tests/test_gpu_kernels.py::test_eye_roll_branches_bit_exact tests/test_gpu_kernels.py::test_match_against_numpy"
# the plain run must be green first (never run a sanitizer on a program that faults)
python -m pytest $TESTS -q -x -k "not impl1 and not 1-" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed" | tee $OUT; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
EXTRA=""
[ "$TOOL" = "racecheck" ] && EXTRA="--racecheck-report all"
timeout 1500 compute-sanitizer --tool $TOOL $EXTRA --target-processes all --error-exitcode 9 --print-limit 40 \
    python -m pytest $TESTS -q -x -k "0] or bit_exact or numpy" > $OUT 2>&1
echo "exit code $?" >> $OUT
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" $OUT | tail -8
