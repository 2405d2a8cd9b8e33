#!/bin/bash
# GPU call 15: launch-lean batch norm (tests + bench), pair-chunk sweep upwards, ncu of the small 256-channel layers.
set -u
OUT=gpurun_out/r2_run15
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== batch-norm tests" | tee "$OUT/summary.txt"
timeout 300 $PT tests/test_gpu_batchnorm.py tests/test_gpu_zz_syncbn.py 2>&1 | tail -8 | tee -a "$OUT/summary.txt"
echo "== full GPU suite" | tee -a "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -5 | tee -a "$OUT/summary.txt"
echo "== bench cfg3 (x2)" | tee -a "$OUT/summary.txt"
for v in a b; do
  timeout 400 python bench.py --no-cpu-baseline 2>"$OUT/bench_${v}_err.txt" | tail -1 > "$OUT/bench_$v.json"
  python -c "import json; d=json.load(open('$OUT/bench_$v.json')); print(round(d['ms_per_step'],2),'ms/step; e2e', round(d['e2e']['ms_per_step'],2), 'ms; launches', d['gpu_launches'])" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_${v}_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
done
echo "== step profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/step_profile.py > "$OUT/step_profile.txt" 2>&1
grep -v Warn "$OUT/step_profile.txt" | head -26 | cut -c1-150 | tee -a "$OUT/summary.txt"
echo "== pair chunks 131072 / 262144 / 1048576 rows (wgrad column)" | tee -a "$OUT/summary.txt"
for cr in 131072 262144 1048576; do
  MEB200_PAIR_CHUNK_ROWS=$cr timeout 200 python profiles/layer_bench.py --only "block" 2>&1 | grep "block1 \|block7 96\|block8" | cut -c1-200 | sed "s/^/[$cr] /" | tee -a "$OUT/summary.txt"
done
echo "== ncu: small layers (block4 256->256, 3.5k rows)" | tee -a "$OUT/summary.txt"
for kn in k_wgrad_pairs k_conv_ts; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -o "$OUT/ncu_b4_$kn" -f \
    python profiles/layer_bench.py --only "block4" --iters 1 > "$OUT/ncu_b4_$kn.log" 2>&1
  tail -1 "$OUT/ncu_b4_$kn.log" | cut -c1-160 | tee -a "$OUT/summary.txt"
done
