#!/bin/bash
# GPU call 25 (last of the round): suite on the final sources (batched weight packing on), bench A/B.
set -u
OUT=gpurun_out/r2_run25
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 120 --timeout-method=thread"
echo "== full GPU suite" | tee "$OUT/summary.txt"
timeout 200 $PT tests -m gpu 2>&1 | tail -12 | cut -c1-220 | tee -a "$OUT/summary.txt"
b() {  # name env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --no-cpu-baseline 2>"$OUT/bench_${name}_err.txt" | tail -1 > "$OUT/bench_$name.json"
  python -c "import json; d=json.load(open('$OUT/bench_$name.json')); print('$name', round(d['ms_per_step'],2),'ms/step; e2e', round(d['e2e']['ms_per_step'],2), 'ms; launches', d['gpu_launches'], d.get('loss_check',{}).get('ok'))" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_${name}_err.txt" | tail -6 | tee -a "$OUT/summary.txt"
}
echo "== bench cfg3: batched packing on / off" | tee -a "$OUT/summary.txt"
b batched A=1
b single MEB200_PACK_BATCHED=0
