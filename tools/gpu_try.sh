#!/bin/bash
# Development visit of the GPU box: the new tcgen05 kernels in isolation, one pytest process per variant (a trap in one variant
# poisons only its own CUDA context), then the kernel timing table.  Output under gpurun_out/<tag>_*.
set -u
TAG=${1:-try}
OUT=gpurun_out
mkdir -p "$OUT"
run() {  # name, timeout, pytest -k expression
    timeout "$2" python -m pytest tests/test_gpu_ops.py -q -m gpu -x -p no:cacheprovider -k "$3" > "$OUT/${TAG}_$1.log" 2>&1
    echo "$1 rc=$? : $(tail -1 "$OUT/${TAG}_$1.log")"
}
run single_cta 300 "test_encoder_forward_backward_vs_oracle and (tc-single-cta or fp32)"
run fused_single 300 "fused and single"
run pair 300 "test_encoder_forward_backward_vs_oracle and (tc or tc-3prod) and not single"
run fused_pair 300 "fused and pair"
run cache 300 "operand_cache or philox"
timeout 600 python tools/time_enc.py 4096 200 5 > "$OUT/${TAG}_time_enc.json" 2> "$OUT/${TAG}_time_enc.err"
echo "time_enc rc=$?"; cat "$OUT/${TAG}_time_enc.json"; tail -3 "$OUT/${TAG}_time_enc.err"
