#!/bin/bash
# 2-GPU call: SyncBN numerics (NCCL exchange and NVLink peer-memory exchange), N=2 bench both ways.
set -u
OUT=gpurun_out/r2_run11
mkdir -p "$OUT"
PT="python -m pytest -q --timeout 300 --timeout-method=thread"
echo "== batch-norm kernels (unrolled passes)" | tee "$OUT/summary.txt"
timeout 300 $PT tests/test_gpu_batchnorm.py 2>&1 | tail -3 | tee -a "$OUT/summary.txt"
echo "== SyncBN numerics, NCCL exchange (1 and 2 ranks)" | tee -a "$OUT/summary.txt"
timeout 400 $PT tests/test_gpu_zz_syncbn.py 2>&1 | tail -6 | tee -a "$OUT/summary.txt"
echo "== SyncBN numerics, peer-memory exchange" | tee -a "$OUT/summary.txt"
MEB200_SYNCBN_PEER=1 timeout 400 $PT tests/test_gpu_zz_syncbn.py 2>&1 | tail -12 | tee -a "$OUT/summary.txt"
run2() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 \
     bench.py --gpus 2 --steps 8 --warmup 3 2>"$OUT/bench_n2_$2.err" | tail -1 > "$OUT/bench_n2_$2.json"
  cut -c1-330 "$OUT/bench_n2_$2.json" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback" -A4 "$OUT/bench_n2_$2.err" | tail -10 | tee -a "$OUT/summary.txt"
}
echo "== bench N=2, NCCL SyncBN exchange" | tee -a "$OUT/summary.txt"
run2 29611 nccl
echo "== bench N=2, peer-memory SyncBN exchange" | tee -a "$OUT/summary.txt"
MEB200_SYNCBN_PEER=1 run2 29612 peer
echo "== bench N=2, plain DDP (reference example settings)" | tee -a "$OUT/summary.txt"
MEB200_DDP_PLAIN=1 run2 29613 plain
echo "== bench N=1 on the same box" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-330 | tee -a "$OUT/summary.txt"
