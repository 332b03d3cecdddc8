#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): data-parallel parity tests, then the bench at N ranks (carries dp_check).
set -u
TAG=${1:-dp}
N=${2:-2}
OUT=gpurun_out
mkdir -p "$OUT"
timeout 1500 python -m pytest tests/test_dp.py -q -m gpu -p no:cacheprovider > "$OUT/${TAG}_dp_tests.log" 2>&1
echo "dp tests rc=$? : $(tail -1 "$OUT/${TAG}_dp_tests.log")"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 \
    > "$OUT/${TAG}_bench_n$N.json" 2> "$OUT/${TAG}_bench_n$N.err"
echo "bench n=$N rc=$?"; cut -c1-900 "$OUT/${TAG}_bench_n$N.json"; python3 -c "
import json,sys
d=json.load(open('$OUT/${TAG}_bench_n$N.json')); print('dp_check', d.get('dp_check')); print('e2e', d.get('e2e')); print('ms', d['ms_per_step'])" ; tail -5 "$OUT/${TAG}_bench_n$N.err"
