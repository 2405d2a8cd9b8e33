#!/bin/bash
# GPU call 24: bench A/B of the NUMA binding of the bench ranks (1 GPU), on the final sources.
set -u
OUT=gpurun_out/r2_run24
mkdir -p "$OUT"
b() {  # name env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline 2>"$OUT/bench_${name}_err.txt" | tail -1 > "$OUT/bench_$name.json"
  python -c "import json; d=json.load(open('$OUT/bench_$name.json')); print('$name', round(d['ms_per_step'],2),'ms/step; e2e', round(d['e2e']['ms_per_step'],2), 'ms; cpus bound', d['config'].get('host_cpus_bound'))" 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_${name}_err.txt" | tail -6 | tee -a "$OUT/summary.txt"
}
echo "== bench cfg3: NUMA binding on / off / on / off" | tee "$OUT/summary.txt"
b bind1 A=1
b nobind1 MEB200_BENCH_AFFINITY=0
b bind2 A=1
b nobind2 MEB200_BENCH_AFFINITY=0
nproc | tee -a "$OUT/summary.txt"; uptime | tee -a "$OUT/summary.txt"
