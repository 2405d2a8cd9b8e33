#!/bin/bash
# GPU call 17: stem temporaries fixed; kernel maps prefetched behind the input map.
set -u
OUT=gpurun_out/r2_run17
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== stem + tc tests" | tee "$OUT/summary.txt"
timeout 600 $PT tests/test_gpu_stem.py tests/test_gpu_prefetch.py tests/test_gpu_tc.py 2>&1 | tail -15 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== layers" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== full GPU suite" | tee -a "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -5 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== bench cfg3 (x2), then with MEB200_STEM_TC=0" | tee -a "$OUT/summary.txt"
for v in a b; do
  timeout 400 python bench.py --no-cpu-baseline 2>"$OUT/bench_${v}_err.txt" | tail -1 > "$OUT/bench_$v.json"
  python -c "import json; d=json.load(open('$OUT/bench_$v.json')); print(round(d['ms_per_step'],2),'ms/step; e2e', round(d['e2e']['ms_per_step'],2), 'ms; launches', d['gpu_launches'], d.get('loss_check'))" 2>&1 | tail -1 | cut -c1-300 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_${v}_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
done
MEB200_MAP_PREFETCH=0 timeout 400 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no map prefetch:', round(d['ms_per_step'],2),'ms/step')" | tee -a "$OUT/summary.txt"
echo "== step profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/step_profile.py > "$OUT/step_profile.txt" 2>&1
grep -v Warn "$OUT/step_profile.txt" | head -30 | cut -c1-150 | tee -a "$OUT/summary.txt"
echo "== host profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/cpu_profile.py > "$OUT/cpu_profile.txt" 2>&1
head -3 "$OUT/cpu_profile.txt" | cut -c1-170 | tee -a "$OUT/summary.txt"
