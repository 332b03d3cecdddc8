#!/bin/bash
# Development visit: op-level regression (forward + backward variants), timing table, full GPU suite, default bench.
set -u
TAG=${1:-try}
OUT=gpurun_out
mkdir -p "$OUT"
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider > "$OUT/${TAG}_ops.log" 2>&1
echo "ops rc=$? : $(tail -1 "$OUT/${TAG}_ops.log")"
timeout 600 python tools/time_enc.py 4096 200 12 > "$OUT/${TAG}_time_enc.json" 2> "$OUT/${TAG}_time_enc.err"
echo "time_enc rc=$?"; cat "$OUT/${TAG}_time_enc.json"; tail -3 "$OUT/${TAG}_time_enc.err"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_ops.py > "$OUT/${TAG}_gpu_tests.log" 2>&1
echo "gpu tests rc=$? : $(tail -1 "$OUT/${TAG}_gpu_tests.log")"
timeout 600 python bench.py > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"
echo "bench rc=$?"; cut -c1-1500 "$OUT/${TAG}_bench.json"; tail -5 "$OUT/${TAG}_bench.err"
# per-launch durations of one eager step (no graph so that every kernel is listed; numbers under ncu are never bench values)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > "$OUT/${TAG}_ncu_launches.log" 2>&1
echo "launch list rc=$?"
