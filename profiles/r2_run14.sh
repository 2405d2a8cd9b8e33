#!/bin/bash
# GPU call 14: stride-pyramid prefetch (tests, bench on/off), pair chunk 131072, ncu capture of k_kernel_map (cfg4).
set -u
OUT=gpurun_out/r2_run14
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== prefetch tests" | tee "$OUT/summary.txt"
timeout 300 $PT tests/test_gpu_prefetch.py 2>&1 | tail -12 | tee -a "$OUT/summary.txt"
echo "== full GPU suite" | tee -a "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -6 | tee -a "$OUT/summary.txt"
echo "== bench cfg3: prefetch on (default), off, on" | tee -a "$OUT/summary.txt"
for v in 1 0 1; do
  MEB200_MAP_PREFETCH=$v timeout 400 python bench.py --no-cpu-baseline 2>"$OUT/bench_pf${v}_err.txt" | tail -1 > "$OUT/bench_pf$v.json"
  python -c "import json; d=json.load(open('$OUT/bench_pf$v.json')); print('prefetch=$v', round(d['ms_per_step'],2),'ms/step; e2e', round(d['e2e']['ms_per_step'],2), 'ms; launches', d['gpu_launches'])" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_pf${v}_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
done
echo "== bench cfg1, cfg2, cfg4" | tee -a "$OUT/summary.txt"
for c in cfg1 cfg2 cfg4; do
  timeout 400 python bench.py --config $c --no-cpu-baseline 2>"$OUT/bench_${c}_err.txt" | tail -1 > "$OUT/bench_$c.json"
  python -c "import json; d=json.load(open('$OUT/bench_$c.json')); print('$c', round(d['ms_per_step'],3),'ms/step; e2e', round(d['e2e']['ms_per_step'],3), 'ms; roofline', d['roofline'].get('bound'), round(d['roofline'].get('frac',0),3))" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
done
echo "== host profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/cpu_profile.py > "$OUT/cpu_profile.txt" 2>&1
head -3 "$OUT/cpu_profile.txt" | cut -c1-170 | tee -a "$OUT/summary.txt"
echo "== pair chunks of 131072 rows" | tee -a "$OUT/summary.txt"
MEB200_PAIR_CHUNK_ROWS=131072 timeout 200 python profiles/layer_bench.py --only "block8" 2>&1 | tail -3 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== ncu: k_kernel_map (cfg4, K=81, 4-D)" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_kernel_map -c 1 -o "$OUT/ncu_k_kernel_map" -f \
    python bench.py --config cfg4 --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_k_kernel_map.log" 2>&1
tail -1 "$OUT/ncu_k_kernel_map.log" | cut -c1-200 | tee -a "$OUT/summary.txt"
