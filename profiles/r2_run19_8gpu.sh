#!/bin/bash
# 8-GPU call: weak scaling of the default path (DDP + SyncBN over peer memory) and its limiter.
set -u
OUT=gpurun_out/r2_run19
mkdir -p "$OUT"
P=29700
runN() {   # name, n, env...
  name=$1; n=$2; shift 2
  P=$((P+1))
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
     bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline 2>"$OUT/bench_$name.err" | tail -1 > "$OUT/bench_$name.json"
  echo "-- $name: $(python -c "import json,sys; d=json.load(open('$OUT/bench_$name.json')); print('n', d['n_gpus'], round(d['ms_per_step'],2),'ms/step', round(d['value']/1e6,2),'Mvox/s; e2e', round(d['e2e']['ms_per_step'],2), 'ms; clocks', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>&1 | tail -1)" | tee -a "$OUT/summary.txt"
  grep -a "Error\|error\|Traceback\|Warning: MEB200\|symmetric" -A3 "$OUT/bench_$name.err" | tail -8 | tee -a "$OUT/summary.txt"
}
echo "== weak scaling, 8 clouds x 100k voxels per rank" | tee "$OUT/summary.txt"
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > "$OUT/bench_n1.json"
python -c "import json; d=json.load(open('$OUT/bench_n1.json')); print('-- n1:', round(d['ms_per_step'],2),'ms/step', round(d['value']/1e6,2),'Mvox/s')" | tee -a "$OUT/summary.txt"
runN n8 8 A=1
runN n8_nosyncbn 8 MEB200_BENCH_ABLATE=nosyncbn
runN n8_nccl 8 MEB200_SYNCBN_PEER=0
runN n2 2 A=1
