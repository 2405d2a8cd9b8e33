#!/bin/bash
# GPU call 5: cp.async index rings (k_conv_ts, k_wgrad_pairs), chunked pair lists.
set -u
OUT=gpurun_out/r2_run5
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== tensor-core + quantize tests" | tee "$OUT/summary.txt"
timeout 500 $PT tests/test_gpu_tc.py tests/test_gpu_quantize.py 2>&1 | tail -12 | tee -a "$OUT/summary.txt"
echo "== layers, defaults" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== layers, ts R=1 acc=2" | tee -a "$OUT/summary.txt"
MEB200_TA_R=1 MEB200_TA_ACC=2 timeout 200 python profiles/layer_bench.py 2>&1 | tail -14 | cut -c1-150 | tee -a "$OUT/summary.txt"
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 400 python bench.py 2>"$OUT/bench_err.txt" | tail -1 > "$OUT/bench_cfg3.json"
cut -c1-600 "$OUT/bench_cfg3.json" | tee -a "$OUT/summary.txt"
grep -a "Error\|error\|Traceback" -A3 "$OUT/bench_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
echo "== ncu captures (block8 96->96)" | tee -a "$OUT/summary.txt"
for kn in k_conv_ts k_wgrad_pairs; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -o "$OUT/ncu_$kn" -f \
    python profiles/layer_bench.py --only "block8 96" --iters 1 > "$OUT/ncu_$kn.log" 2>&1
  tail -1 "$OUT/ncu_$kn.log" | tee -a "$OUT/summary.txt"
done
