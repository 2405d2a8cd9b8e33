#!/bin/bash
# GPU call 22: cfg2 (small, host-bound network) A/B of the round's host-side changes; suite on the final Python.
set -u
OUT=gpurun_out/r2_run22
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== full GPU suite" | tee "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -3 | cut -c1-200 | tee -a "$OUT/summary.txt"
b() {  # name cfg env...
  name=$1; cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --no-cpu-baseline 2>"$OUT/bench_${name}_err.txt" | tail -1 > "$OUT/bench_$name.json"
  python -c "import json; d=json.load(open('$OUT/bench_$name.json')); print('$name', round(d['ms_per_step'],3),'ms/step; e2e', round(d['e2e']['ms_per_step'],3), 'ms; launches', d['gpu_launches'])" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_${name}_err.txt" | tail -6 | tee -a "$OUT/summary.txt"
}
echo "== cfg2 A/B" | tee -a "$OUT/summary.txt"
b cfg2_default cfg2 A=1
b cfg2_noprefetch cfg2 MEB200_MAP_PREFETCH=0
b cfg2_stem_simt cfg2 MEB200_STEM_TC=0
b cfg2_default2 cfg2 A=1
echo "== cfg4 / cfg1 A/B" | tee -a "$OUT/summary.txt"
b cfg4_default cfg4 A=1
b cfg4_noprefetch cfg4 MEB200_MAP_PREFETCH=0
b cfg1_default cfg1 A=1
b cfg1_noprefetch cfg1 MEB200_MAP_PREFETCH=0
