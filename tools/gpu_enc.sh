#!/bin/bash
# Encoder-kernel development visit: op-level parity tests, then the timing table of every forward / backward variant.
set -u
TAG=${1:-enc}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p "$OUT"
if [ -n "$KEXPR" ]; then
  timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider -k "$KEXPR" > "$OUT/${TAG}_ops.log" 2>&1
else
  timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider > "$OUT/${TAG}_ops.log" 2>&1
fi
rc=$?
echo "ops rc=$rc : $(tail -1 "$OUT/${TAG}_ops.log")"
[ $rc -ne 0 ] && { tail -40 "$OUT/${TAG}_ops.log"; exit 0; }
timeout 600 python tools/time_enc.py 4096 200 10 > "$OUT/${TAG}_time_enc.json" 2> "$OUT/${TAG}_time_enc.err"
echo "time_enc rc=$?"; cat "$OUT/${TAG}_time_enc.json"; tail -3 "$OUT/${TAG}_time_enc.err"
