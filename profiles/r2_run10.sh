#!/bin/bash
# GPU call 10: full suite + bench lines + ncu captures of the final kernels.
set -u
OUT=gpurun_out/r2_run10
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== full GPU suite" | tee "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -12 | tee -a "$OUT/summary.txt"
echo "== layers, defaults" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== bench lines" | tee -a "$OUT/summary.txt"
for c in cfg3 cfg1 cfg2 cfg4; do
  timeout 400 python bench.py --config $c 2>"$OUT/bench_${c}_err.txt" | tail -1 > "$OUT/bench_$c.json"
  cut -c1-420 "$OUT/bench_$c.json" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A3 "$OUT/bench_${c}_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
done
echo "== step profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/step_profile.py > "$OUT/step_profile.txt" 2>&1
head -40 "$OUT/step_profile.txt" | cut -c1-160 | tee -a "$OUT/summary.txt"
echo "== ncu captures (block8 96->96)" | tee -a "$OUT/summary.txt"
for kn in k_conv_ts k_wgrad_pairs; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -o "$OUT/ncu_$kn" -f \
    python profiles/layer_bench.py --only "block8 96" --iters 1 > "$OUT/ncu_$kn.log" 2>&1
  tail -1 "$OUT/ncu_$kn.log" | tee -a "$OUT/summary.txt"
done
