#!/bin/bash
# 2-GPU call: fused peer exchange (default) + where the N=2 step loses time (ablations).
set -u
OUT=gpurun_out/r2_run12
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 300 --timeout-method=thread"
echo "== SyncBN numerics, peer-memory exchange (default)" | tee "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_zz_syncbn.py 2>&1 | tail -8 | tee -a "$OUT/summary.txt"
echo "== SyncBN numerics, NCCL exchange" | tee -a "$OUT/summary.txt"
MEB200_SYNCBN_PEER=0 timeout 400 $PT tests/test_gpu_zz_syncbn.py 2>&1 | tail -4 | tee -a "$OUT/summary.txt"
P=29620
run2() {   # name, env...
  name=$1; shift
  P=$((P+1))
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P \
     bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>"$OUT/bench_n2_$name.err" | tail -1 > "$OUT/bench_n2_$name.json"
  echo "-- $name: $(python -c "import json,sys; d=json.load(open('$OUT/bench_n2_$name.json')); print(round(d['ms_per_step'],2),'ms/step', round(d['value']/1e6,2),'Mvox/s e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])" 2>&1 | tail -1)" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_n2_$name.err" | tail -8 | tee -a "$OUT/summary.txt"
}
echo "== bench N=2" | tee -a "$OUT/summary.txt"
run2 default A=1
run2 nccl MEB200_SYNCBN_PEER=0
run2 nosyncbn MEB200_BENCH_ABLATE=nosyncbn
run2 noddp MEB200_BENCH_ABLATE=noddp
run2 replicas MEB200_BENCH_ABLATE=nosyncbn,noddp
run2 samedata MEB200_BENCH_ABLATE=samedata
run2 default_again A=1
echo "== bench N=1 on the same box" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-330 | tee -a "$OUT/summary.txt"
