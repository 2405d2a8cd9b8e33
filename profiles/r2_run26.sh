#!/bin/bash
# GPU call 26: does the one failure of r2_run25 (kernel-map pair count) follow the batched weight packing?
set -u
OUT=gpurun_out/r2_run26
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 100 --timeout-method=thread -p no:cacheprovider"
echo "== suite, MEB200_PACK_BATCHED=0" | tee "$OUT/summary.txt"
MEB200_PACK_BATCHED=0 timeout 100 $PT tests -m gpu --deselect tests/test_gpu_tc.py::test_batched_weight_packing_equals_single 2>&1 | grep -a "passed\|failed\|FAILED\|^E " | head -12 | cut -c1-220 | tee -a "$OUT/summary.txt"
echo "== suite, default" | tee -a "$OUT/summary.txt"
timeout 100 $PT tests -m gpu 2>&1 | grep -a "passed\|failed\|FAILED\|^E " | head -12 | cut -c1-220 | tee -a "$OUT/summary.txt"
