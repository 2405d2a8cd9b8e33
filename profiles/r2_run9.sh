#!/bin/bash
# GPU call 9: k_wgrad_pairs templated on the slot count.
set -u
OUT=gpurun_out/r2_run9
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== tensor-core tests" | tee "$OUT/summary.txt"
timeout 500 $PT tests/test_gpu_tc.py tests/test_gpu_large.py -x 2>&1 | tail -12 | tee -a "$OUT/summary.txt"
echo "== layers, defaults" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== ncu wgrad" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_pairs -c 1 -o "$OUT/ncu_k_wgrad_pairs" -f python profiles/layer_bench.py --only "block8 96" --iters 1 > "$OUT/ncu_k_wgrad_pairs.log" 2>&1
tail -1 "$OUT/ncu_k_wgrad_pairs.log" | tee -a "$OUT/summary.txt"
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 400 python bench.py 2>"$OUT/bench_err.txt" | tail -1 > "$OUT/bench_cfg3.json"
cut -c1-400 "$OUT/bench_cfg3.json" | tee -a "$OUT/summary.txt"
timeout 400 python bench.py --config cfg4 2>>"$OUT/bench_err.txt" | tail -1 > "$OUT/bench_cfg4.json"
cut -c1-400 "$OUT/bench_cfg4.json" | tee -a "$OUT/summary.txt"
grep -a "Error\|error\|Traceback" -A3 "$OUT/bench_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
