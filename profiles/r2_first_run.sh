#!/bin/bash
# First GPU call of round 2: run every experiment that was prepared (but not measured) at the
# end of round 1, each behind its environment switch, and leave the numbers in gpurun_out/.
#
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash profiles/r2_first_run.sh'   (about 12-15 GPU-minutes)
#
# Every step runs under `timeout`; a failing variant does not stop the others.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
PT="python -m pytest -x -q --timeout 120 --timeout-method=thread"

echo "== defaults: full GPU suite" | tee "$OUT/summary.txt"
timeout 400 python -m pytest -q --timeout 120 --timeout-method=thread tests -m gpu 2>&1 | tail -15 | tee -a "$OUT/summary.txt"
timeout 120 python profiles/layer_bench.py --json "$OUT/layers_default.json" > "$OUT/layers_default.txt" 2>&1
tail -13 "$OUT/layers_default.txt" | cut -c1-200 >> "$OUT/summary.txt"

echo "== SyncBN numerics, one rank (NCCL exchange, then peer-memory exchange)" | tee -a "$OUT/summary.txt"
MEB200_SYNCBN_PEER=1 timeout 300 $PT tests/test_gpu_zz_syncbn.py 2>&1 | tail -12 | tee -a "$OUT/summary.txt"

for PF in 8 16; do
  echo "== wgrad index prefetch PF=$PF" | tee -a "$OUT/summary.txt"
  MEB200_WG_PREFETCH=$PF timeout 200 $PT tests/test_gpu_tc.py tests/test_gpu_large.py -k "wgrad or cfg1" 2>&1 | tail -2 | tee -a "$OUT/summary.txt"
  MEB200_WG_PREFETCH=$PF timeout 120 python profiles/layer_bench.py --json "$OUT/layers_wg_pf$PF.json" > "$OUT/layers_wg_pf$PF.txt" 2>&1
  tail -13 "$OUT/layers_wg_pf$PF.txt" | cut -c150-215 >> "$OUT/summary.txt"
done

echo "== cp.async forward producers with the 4-deep index ring" | tee -a "$OUT/summary.txt"
MEB200_TC_GATHER=cpasync timeout 120 python profiles/layer_bench.py --json "$OUT/layers_cpasync.json" > "$OUT/layers_cpasync.txt" 2>&1
tail -13 "$OUT/layers_cpasync.txt" | cut -c1-150 >> "$OUT/summary.txt"

for E in 1 2 6 7; do
  echo "== k_conv_rs experiment MEB200_RS_EXP=$E (96-channel bf16 shape only)" | tee -a "$OUT/summary.txt"
  MEB200_RS_EXP=$E timeout 200 $PT tests/test_gpu_tc.py -k "forward or layer" 2>&1 | tail -2 | tee -a "$OUT/summary.txt"
  MEB200_RS_EXP=$E timeout 120 python profiles/layer_bench.py --only "96" --json "$OUT/layers_rs_exp$E.json" > "$OUT/layers_rs_exp$E.txt" 2>&1
  tail -5 "$OUT/layers_rs_exp$E.txt" | cut -c1-150 >> "$OUT/summary.txt"
done

echo "== ordered (Morton) tiles + L1-allocating gathers, forward of the 96-channel layers" | tee -a "$OUT/summary.txt"
for KB in 0 144; do
  for E in 0 2; do
    echo "-- MEB200_TC_SMEM_KB=$KB MEB200_RS_EXP=$E" >> "$OUT/summary.txt"
    MEB200_TC_SMEM_KB=$KB MEB200_RS_EXP=$E timeout 120 python profiles/layer_bench.py --only "96->96" --ordered \
      > "$OUT/layers_ordered_kb${KB}_e$E.txt" 2>&1
    grep -a "ordered\|96->96" "$OUT/layers_ordered_kb${KB}_e$E.txt" | cut -c1-150 >> "$OUT/summary.txt"
  done
done

echo "== operand A in tensor memory: layout micro-test" | tee -a "$OUT/summary.txt"
(cd profiles/micro && timeout 30 ./tmem_a_layout) 2>&1 | tee -a "$OUT/summary.txt"

echo "== bench with defaults" | tee -a "$OUT/summary.txt"
timeout 200 python bench.py 2>/dev/null | tail -1 > "$OUT/bench_default.json"
cut -c1-330 "$OUT/bench_default.json" | tee -a "$OUT/summary.txt"
