#!/bin/bash
# GPU call 4 of round 2: deep index prefetch, ncu captures of the new kernels, new tests, cfg lines.
set -u
OUT=gpurun_out/r2_run4
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== tensor-core tests" | tee "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_tc.py 2>&1 | tail -5 | tee -a "$OUT/summary.txt"
echo "== layers, defaults" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== layers, ts R=1 acc=2" | tee -a "$OUT/summary.txt"
MEB200_TA_R=1 MEB200_TA_ACC=2 timeout 200 python profiles/layer_bench.py 2>&1 | tail -14 | cut -c1-150 | tee -a "$OUT/summary.txt"
echo "== new tests: quantize, batch norm (fused), network bf16" | tee -a "$OUT/summary.txt"
timeout 600 $PT tests/test_gpu_quantize.py tests/test_gpu_batchnorm.py tests/test_gpu_network.py -s 2>&1 | grep -a "bf16 vs reference\|grad \|passed\|failed\|Error\|error\|FAILED\|assert" | head -60 | tee -a "$OUT/summary.txt"
echo "== bench default + cfg1 cfg2 cfg4" | tee -a "$OUT/summary.txt"
for c in cfg3 cfg1 cfg2 cfg4; do
  timeout 400 python bench.py --config $c 2>"$OUT/bench_${c}_err.txt" | tail -1 > "$OUT/bench_$c.json"
  cut -c1-1800 "$OUT/bench_$c.json" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A3 "$OUT/bench_${c}_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
done
echo "== ncu captures (block8 96->96): k_conv_ts, k_wgrad_pairs" | tee -a "$OUT/summary.txt"
for kn in k_conv_ts k_wgrad_pairs; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -o "$OUT/ncu_$kn" -f \
    python profiles/layer_bench.py --only "block8 96" --iters 1 > "$OUT/ncu_$kn.log" 2>&1
  tail -2 "$OUT/ncu_$kn.log" | tee -a "$OUT/summary.txt"
done
ls -la "$OUT" | tee -a "$OUT/summary.txt"
