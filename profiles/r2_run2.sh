#!/bin/bash
# GPU call 2 of round 2: k_conv_ta (operand A in tensor memory) correctness + per-layer timing.
set -u
OUT=gpurun_out/r2_run2
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== k_conv_ta sweep" | tee "$OUT/summary.txt"
timeout 300 $PT tests/test_gpu_tc.py -x 2>&1 | tail -25 | tee -a "$OUT/summary.txt"
echo "== bf16 network parity (prints measured errors)" | tee -a "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_network.py -k bf16 -s 2>&1 | grep -v Warning | tail -60 | tee -a "$OUT/summary.txt"
echo "== layers, TA default" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers_ta.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
for R in "1 2" "2 1" "4 1"; do
  set -- $R
  echo "== layers, TA R=$1 acc=$2" | tee -a "$OUT/summary.txt"
  MEB200_TA_R=$1 MEB200_TA_ACC=$2 timeout 200 python profiles/layer_bench.py 2>&1 | tail -14 | cut -c1-150 | tee -a "$OUT/summary.txt"
done
echo "== layers, RS (round-1 kernel, packed weights)" | tee -a "$OUT/summary.txt"
MEB200_TC_KERNEL=rs timeout 200 python profiles/layer_bench.py 2>&1 | tail -14 | cut -c1-150 | tee -a "$OUT/summary.txt"
echo "== rest of the GPU suite" | tee -a "$OUT/summary.txt"
timeout 600 $PT tests -m gpu --deselect tests/test_gpu_tc.py -k "not bf16_matches" 2>&1 | tail -8 | tee -a "$OUT/summary.txt"
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py 2>"$OUT/bench_err.txt" | tail -1 > "$OUT/bench_default.json"
cut -c1-1500 "$OUT/bench_default.json" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/bench_err.txt" | tee -a "$OUT/summary.txt"
