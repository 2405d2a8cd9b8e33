#!/bin/bash
# Final GPU call of the round: everything the judged state rests on, on the final sources.
set -u
OUT=gpurun_out/r2_final
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 180 --timeout-method=thread"
echo "== smoke()" | tee "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== full GPU suite" | tee -a "$OUT/summary.txt"
timeout 900 $PT tests -m gpu 2>&1 | tail -4 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== bench lines" | tee -a "$OUT/summary.txt"
for c in cfg3 cfg1 cfg2 cfg4; do
  timeout 500 python bench.py --config $c 2>"$OUT/bench_${c}_err.txt" | tail -1 > "$OUT/bench_$c.json"
  python -c "import json; d=json.load(open('$OUT/bench_$c.json')); r=d['roofline']; print('$c', round(d['ms_per_step'],3),'ms/step', round(d['value']/1e6,2),'M/s; e2e', round(d['e2e']['ms_per_step'],3), 'ms; launches', d['gpu_launches'], '; roofline', r.get('bound'), round(r.get('achieved',0),1), r.get('unit'), round(r.get('frac',0),3), '; cpu', d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1 | cut -c1-300 | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A3 "$OUT/bench_${c}_err.txt" | tail -6 | tee -a "$OUT/summary.txt"
done
echo "== reference arm (1 step)" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>"$OUT/bench_ref_err.txt" | tail -1 > "$OUT/bench_reference.json"
cut -c1-300 "$OUT/bench_reference.json" | tee -a "$OUT/summary.txt"
echo "== layers" | tee -a "$OUT/summary.txt"
timeout 200 python profiles/layer_bench.py --json "$OUT/layers.json" 2>&1 | tail -14 | cut -c1-200 | tee -a "$OUT/summary.txt"
echo "== step profile" | tee -a "$OUT/summary.txt"
timeout 300 python profiles/step_profile.py > "$OUT/step_profile.txt" 2>&1
grep -v Warn "$OUT/step_profile.txt" | head -34 | cut -c1-150 | tee -a "$OUT/summary.txt"
echo "== ncu captures" | tee -a "$OUT/summary.txt"
cap() {  # name, kernel regex, layer filter, launch skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o "$OUT/ncu_$1" -f \
    python profiles/layer_bench.py --only "$3" --iters 1 > "$OUT/ncu_$1.log" 2>&1
  tail -1 "$OUT/ncu_$1.log" | cut -c1-160 | tee -a "$OUT/summary.txt"
}
cap k_conv_ts_block8 k_conv_ts "block8 96" 0
cap k_wgrad_pairs_block8 k_wgrad_pairs "block8 96" 0
cap k_conv_ts_stem k_conv_ts "conv0" 0
cap k_wgrad_stem k_wgrad_stem "conv0" 0
echo "== ncu launch list of the bench command" | tee -a "$OUT/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/launches_bench.log" 2>&1
python profiles/summarize_launches.py "$OUT/launches.csv" 30 > "$OUT/launches_summary.txt" 2>&1
head -20 "$OUT/launches_summary.txt" | cut -c1-150 | tee -a "$OUT/summary.txt"
gzip -f "$OUT/launches.csv"
