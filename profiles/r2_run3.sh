#!/bin/bash
# GPU call 3 of round 2: k_conv_ts (staged A-in-TMEM), k_wgrad_pairs, pair lists.
set -u
OUT=gpurun_out/r2_run3
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== tensor-core tests (ts + pairs default)" | tee "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_tc.py 2>&1 | tail -30 | tee -a "$OUT/summary.txt"
echo "== v1 kernel still green (MEB200_TC_KERNEL=ta)" | tee -a "$OUT/summary.txt"
MEB200_TC_KERNEL=ta timeout 300 $PT tests/test_gpu_tc.py -k "ta_forward or layer" 2>&1 | tail -5 | tee -a "$OUT/summary.txt"
echo "== bf16 network parity (prints measured errors)" | tee -a "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_network.py -k bf16 -s 2>&1 | grep -a "bf16 vs reference\|grad \|passed\|failed\|Error\|error" | tee -a "$OUT/summary.txt"
echo "== layers, defaults (ts + pairs)" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers_ts.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
for R in "1 2" "4 1"; do
  set -- $R
  echo "== layers, ts R=$1 acc=$2 (fwd/dgrad columns)" | tee -a "$OUT/summary.txt"
  MEB200_TA_R=$1 MEB200_TA_ACC=$2 timeout 200 python profiles/layer_bench.py 2>&1 | tail -14 | cut -c1-150 | tee -a "$OUT/summary.txt"
done
echo "== rest of the GPU suite" | tee -a "$OUT/summary.txt"
timeout 900 $PT tests -m gpu --deselect tests/test_gpu_tc.py -k "not bf16_matches" 2>&1 | tail -8 | tee -a "$OUT/summary.txt"
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py 2>"$OUT/bench_err.txt" | tail -1 > "$OUT/bench_default.json"
cut -c1-2500 "$OUT/bench_default.json" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/bench_err.txt" | tee -a "$OUT/summary.txt"
