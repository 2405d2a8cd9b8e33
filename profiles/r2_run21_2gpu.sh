#!/bin/bash
# 2-GPU call: where DDP's 1.5 ms go (bucket size, NCCL CTA count); SyncBN tests after the num_batches_tracked fold.
set -u
OUT=gpurun_out/r2_run21
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 200 --timeout-method=thread"
echo "== SyncBN + BN tests" | tee "$OUT/summary.txt"
timeout 300 $PT tests/test_gpu_zz_syncbn.py tests/test_gpu_batchnorm.py 2>&1 | tail -4 | cut -c1-200 | tee -a "$OUT/summary.txt"
P=29900
run2() {   # name, env...
  name=$1; shift
  P=$((P+1))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P \
     bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>"$OUT/bench_n2_$name.err" | tail -1 > "$OUT/bench_n2_$name.json"
  echo "-- $name: $(python -c "import json,sys; d=json.load(open('$OUT/bench_n2_$name.json')); print(round(d['ms_per_step'],2),'ms/step e2e', round(d['e2e']['ms_per_step'],2))" 2>&1 | tail -1)" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_n2_$name.err" | tail -6 | tee -a "$OUT/summary.txt"
}
echo "== bench N=2" | tee -a "$OUT/summary.txt"
run2 default A=1
run2 bucket200 MEB200_DDP_BUCKET_MB=200
run2 ctas4 NCCL_MAX_CTAS=4
run2 bucket200_ctas4 MEB200_DDP_BUCKET_MB=200 NCCL_MAX_CTAS=4
run2 noddp MEB200_BENCH_ABLATE=noddp
run2 default2 A=1
