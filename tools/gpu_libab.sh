#!/bin/bash
# Same-box A/B of prebuilt library variants.  Box-to-box spread under the power cap (+-8 %) is larger than most kernel changes, so
# variants are compared inside ONE gpurun call: build each variant here, copy it to tools/_libs/lib_<name>.so (git-ignored, travels
# with the snapshot), then   gpurun -- 'bash tools/gpu_libab.sh base variant base variant'
P=robust-speech-emotion-recognition-via-dynamic-asymmetric-distillation-in-noisy-environments_b200
cp $P/libdad_sm100a.so /tmp/lib_orig.so
for v in "$@"; do
  cp tools/_libs/lib_$v.so $P/libdad_sm100a.so
  echo -n "variant $v: "
  timeout 300 python tools/time_enc.py 4096 200 10 2>/dev/null | python3 -c "
import json,sys; d=json.load(sys.stdin); print({k:v for k,v in d.items() if 'single' not in k and k not in ('B','T','reps','alg_GB_per_view')})"
done
cp /tmp/lib_orig.so $P/libdad_sm100a.so
