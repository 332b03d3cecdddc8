#!/bin/bash
# Development visit: regression of the forward variants, timing table, then ncu --set full of the forward kernels.
set -u
TAG=${1:-try}
OUT=gpurun_out
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider -k "fused or test_encoder_forward_backward_vs_oracle or operand_cache or philox" > "$OUT/${TAG}_ops.log" 2>&1
echo "ops rc=$? : $(tail -1 "$OUT/${TAG}_ops.log")"
timeout 600 python tools/time_enc.py 4096 200 5 > "$OUT/${TAG}_time_enc.json" 2> "$OUT/${TAG}_time_enc.err"
echo "time_enc rc=$?"; cat "$OUT/${TAG}_time_enc.json"; tail -3 "$OUT/${TAG}_time_enc.err"
W="clean:single,clean,fused:single,fused"
timeout 120 python tools/prof_enc.py 4096 200 $W > "$OUT/${TAG}_prof_plain.log" 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_enc_fwd_tc -s 1 -c 4 -f -o "$OUT/${TAG}_fwd" \
    python tools/prof_enc.py 4096 200 $W > "$OUT/${TAG}_ncu.log" 2>&1
echo "ncu rc=$?"; tail -2 "$OUT/${TAG}_ncu.log"; ls -la "$OUT" | grep "${TAG}_"
