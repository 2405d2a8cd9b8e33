#!/bin/bash
# GPU call 13: full suite on the current sources; 3 vs 4 producer groups in k_conv_ts; pair-list chunk size; host profile.
set -u
OUT=gpurun_out/r2_run13
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== full GPU suite" | tee "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -12 | tee -a "$OUT/summary.txt"
echo "== layers, 3 producer groups (default build)" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers_g3.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== layers, 4 producer groups" | tee -a "$OUT/summary.txt"
MEB200_LIB=libmeb200_g4.so timeout 200 python profiles/layer_bench.py --json "$OUT/layers_g4.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== g4: tensor-core tests" | tee -a "$OUT/summary.txt"
MEB200_LIB=libmeb200_g4.so timeout 600 $PT tests/test_gpu_tc.py tests/test_gpu_network.py 2>&1 | tail -4 | tee -a "$OUT/summary.txt"
echo "== layers, pair chunks of 32768 / 16384 rows (wgrad column)" | tee -a "$OUT/summary.txt"
for cr in 32768 16384; do
  MEB200_PAIR_CHUNK_ROWS=$cr timeout 200 python profiles/layer_bench.py --only "block8" 2>&1 | tail -3 | cut -c1-200 | tee -a "$OUT/summary.txt"
done
echo "== bench cfg3: default, g4" | tee -a "$OUT/summary.txt"
timeout 400 python bench.py --no-cpu-baseline 2>"$OUT/bench_g3_err.txt" | tail -1 > "$OUT/bench_g3.json"; cut -c1-330 "$OUT/bench_g3.json" | tee -a "$OUT/summary.txt"
MEB200_LIB=libmeb200_g4.so timeout 400 python bench.py --no-cpu-baseline 2>"$OUT/bench_g4_err.txt" | tail -1 > "$OUT/bench_g4.json"; cut -c1-330 "$OUT/bench_g4.json" | tee -a "$OUT/summary.txt"
echo "== step profile (default build)" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/step_profile.py > "$OUT/step_profile.txt" 2>&1
head -30 "$OUT/step_profile.txt" | cut -c1-160 | tee -a "$OUT/summary.txt"
echo "== host profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/cpu_profile.py > "$OUT/cpu_profile.txt" 2>&1
head -60 "$OUT/cpu_profile.txt" | cut -c1-170 | tee -a "$OUT/summary.txt"
