#!/bin/bash
# GPU call 18: A/B on one box: map prefetch on/off, stem TC on/off; host profile.
set -u
OUT=gpurun_out/r2_run18
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== prefetch + stem tests" | tee "$OUT/summary.txt"
timeout 600 $PT tests/test_gpu_prefetch.py tests/test_gpu_stem.py 2>&1 | tail -4 | cut -c1-200 | tee -a "$OUT/summary.txt"
b() {  # name env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu-baseline 2>"$OUT/bench_${name}_err.txt" | tail -1 > "$OUT/bench_$name.json"
  python -c "import json; d=json.load(open('$OUT/bench_$name.json')); print('$name', round(d['ms_per_step'],2),'ms/step; e2e', round(d['e2e']['ms_per_step'],2), 'ms; launches', d['gpu_launches'])" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_${name}_err.txt" | tail -8 | tee -a "$OUT/summary.txt"
}
echo "== bench cfg3" | tee -a "$OUT/summary.txt"
b default A=1
b noprefetch MEB200_MAP_PREFETCH=0
b default2 A=1
b stem_simt MEB200_STEM_TC=0
b default3 A=1
echo "== host profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/cpu_profile.py > "$OUT/cpu_profile.txt" 2>&1
head -3 "$OUT/cpu_profile.txt" | cut -c1-170 | tee -a "$OUT/summary.txt"
grep "build)" "$OUT/cpu_profile.txt" | head -2 | cut -c1-150 | tee -a "$OUT/summary.txt"
