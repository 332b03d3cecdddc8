#!/bin/bash
# transposed fused forward: parity tests and timing per variant (DAD_TC_TP_TEAMS=2|3, DAD_TC_TP_NF=64|128)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-nf}; shift
for v in "$@"; do
  export DAD_TC_TP_TEAMS=${v%%:*} DAD_TC_TP_NF=${v##*:}
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider -k "fused_noisy and transposed" > $OUT/${TAG}_ops_$v.log 2>&1
  echo "teams:nf=$v tests rc=$? : $(tail -1 $OUT/${TAG}_ops_$v.log)"
  timeout 300 python tools/time_enc.py 4096 200 10 2>/dev/null | python3 -c "
import json,sys; d=json.load(sys.stdin); print({k:v for k,v in d.items() if 'tp' in k})"
done
