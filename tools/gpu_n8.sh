#!/bin/bash
# N-GPU visit: bench at N ranks with the given exchanges (p2p / nccl), back to back on the same box
set -u
TAG=${1:-n8}; N=${2:-8}; shift 2
OUT=gpurun_out; mkdir -p $OUT
for ex in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 10 --no-e2e --no-dp-check --exchange $ex \
    > "$OUT/${TAG}_bench_n${N}_$ex.json" 2> "$OUT/${TAG}_bench_n${N}_$ex.err"
  echo "exchange=$ex rc=$? $(python3 -c "import json; d=json.load(open('$OUT/${TAG}_bench_n${N}_$ex.json')); print(d['ms_per_step'], d['config']['exchange'], d['clocks']['sm_mhz'])")"
done
