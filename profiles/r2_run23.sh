#!/bin/bash
# GPU call 23: suite + smoke + bench on the final sources (after the C-ABI clean-up and the pair-chunk default).
set -u
OUT=gpurun_out/r2_run23
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== smoke()" | tee "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== full GPU suite" | tee -a "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -3 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== bench cfg3 (with cpu baseline), x2" | tee -a "$OUT/summary.txt"
for v in a b; do
  timeout 500 python bench.py 2>"$OUT/bench_${v}_err.txt" | tail -1 > "$OUT/bench_$v.json"
  python -c "import json; d=json.load(open('$OUT/bench_$v.json')); r=d['roofline']; print(round(d['ms_per_step'],3),'ms/step', round(d['value']/1e6,2),'M/s; e2e', round(d['e2e']['ms_per_step'],3), 'ms; launches', d['gpu_launches'], '; roofline', round(r.get('achieved',0),1), round(r.get('frac',0),3), [ (round(o['achieved'],1), round(o['frac'],3)) for o in r.get('other',[])])" 2>&1 | tail -1 | cut -c1-300 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A3 "$OUT/bench_${v}_err.txt" | tail -6 | tee -a "$OUT/summary.txt"
done
echo "== layers (wgrad with 262144-row chunks)" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
