#!/bin/bash
# 2-GPU call: SyncBN exchange fused into the reduction kernels (numerics, N=2 bench).
set -u
OUT=gpurun_out/r2_run20
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 300 --timeout-method=thread"
echo "== SyncBN numerics, peer-memory exchange (default), then NCCL" | tee "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_zz_syncbn.py tests/test_gpu_batchnorm.py 2>&1 | tail -6 | cut -c1-200 | tee -a "$OUT/summary.txt"
MEB200_SYNCBN_PEER=0 timeout 400 $PT tests/test_gpu_zz_syncbn.py 2>&1 | tail -3 | cut -c1-200 | tee -a "$OUT/summary.txt"
P=29800
run2() {   # name, env...
  name=$1; shift
  P=$((P+1))
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P \
     bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>"$OUT/bench_n2_$name.err" | tail -1 > "$OUT/bench_n2_$name.json"
  echo "-- $name: $(python -c "import json,sys; d=json.load(open('$OUT/bench_n2_$name.json')); print(round(d['ms_per_step'],2),'ms/step', round(d['value']/1e6,2),'Mvox/s e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])" 2>&1 | tail -1)" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_n2_$name.err" | tail -8 | tee -a "$OUT/summary.txt"
}
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('-- n1:', round(d['ms_per_step'],2),'ms/step')" | tee -a "$OUT/summary.txt"
run2 default A=1
run2 nosyncbn MEB200_BENCH_ABLATE=nosyncbn
run2 default2 A=1
