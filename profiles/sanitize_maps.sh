#!/bin/bash
# First GPU call of the next round (profiles/r2_notes.md §13): the one unreproduced failure of
# round 2 (test_kernel_map_symmetry_100k counted 758 pairs too many in 1 of ~15 suite runs).
# Runs the map-building tests in a loop, with and without the ahead-of-time map building, and once
# under compute-sanitizer (memcheck, racecheck, initcheck).
set -u
OUT=gpurun_out/sanitize_maps
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 300 --timeout-method=thread -p no:cacheprovider"
T="tests/test_gpu_golden.py tests/test_gpu_large.py tests/test_gpu_parity.py tests/test_gpu_prefetch.py"
for i in 1 2 3 4 5 6 7 8; do
  $PT $T 2>&1 | grep -a "passed\|failed\|FAILED" | tee -a "$OUT/loop_default.txt"
  MEB200_MAP_PREFETCH=0 $PT $T 2>&1 | grep -a "passed\|failed\|FAILED" | tee -a "$OUT/loop_noprefetch.txt"
done
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 \
    python -m pytest -q -x tests/test_gpu_large.py::test_kernel_map_symmetry_100k tests/test_gpu_prefetch.py \
    > "$OUT/$tool.txt" 2>&1
  tail -5 "$OUT/$tool.txt"
done
