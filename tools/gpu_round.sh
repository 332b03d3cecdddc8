#!/bin/bash
# One GPU-box visit of a round: parity tests, the bench lines of every workload, the reference arm, the ncu launch list and one
# `ncu --set full` capture of the step's dominant kernels.  Run it THROUGH gpurun from the repo root, e.g.
#     gpurun --timeout 900 -- 'bash tools/gpu_round.sh r02'            # everything (about 6 GPU-minutes)
#     gpurun --timeout 300 -- 'bash tools/gpu_round.sh r02 quick'      # tests + default bench + launch list (about 2.5 minutes)
# Everything lands in gpurun_out/<round>_*; copy what should be judged into profiles/ (python tools/launch_table.py,
# python tools/ncu_summary.py <rep> print the summaries).  Numbers printed by a run under ncu are never bench values.
set -u
R=${1:-r02}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p "$OUT"

timeout 300 python -m pytest tests -m gpu -x -q > "$OUT/${R}_gpu_tests.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/${R}_gpu_tests.log"
tail -3 "$OUT/${R}_gpu_tests.log"

timeout 240 python bench.py > "$OUT/${R}_bench_g4096.json" 2> "$OUT/${R}_bench_g4096.err"
echo "bench rc=$?"; cut -c1-400 "$OUT/${R}_bench_g4096.json"

# per-launch durations of one eager step (same command as the bench, two timed steps, no graph so every kernel is listed)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/${R}_launches_g4096.csv" \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > "$OUT/${R}_ncu_launches.log" 2>&1
echo "launch list rc=$?"

[ "$MODE" = quick ] && exit 0

for wl in b64 t1000; do
    timeout 240 python bench.py --workload $wl > "$OUT/${R}_bench_${wl}.json" 2> "$OUT/${R}_bench_${wl}.err"
    echo "bench $wl rc=$?"
done
timeout 240 python bench.py --impl reference > "$OUT/${R}_bench_reference.json" 2> "$OUT/${R}_bench_reference.err"
echo "reference arm rc=$?"

# full captures.  Launch arithmetic of `bench.py --no-graph` (30 CE-only pretraining steps = 1 forward + 1 backward each, then 3
# warm-up steps of 3 forwards + 2 backwards): the first TIMED step's five encoder launches are launches 75..79 of the encoder
# kernels; its three ECDA tensor-core / elementwise kernels come after the 9 of the warm-up steps.
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_enc_(fwd|bwd)_tc" --launch-skip 75 -c 5 \
    -o "$OUT/${R}_full_encoders" -f python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > "$OUT/${R}_ncu_full_encoders.log" 2>&1
echo "ncu full (encoders) rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_gram_tc|k_pz_tc|k_ecda_p" --launch-skip 9 -c 3 \
    -o "$OUT/${R}_full_ecda" -f python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > "$OUT/${R}_ncu_full_ecda.log" 2>&1
echo "ncu full (ECDA) rc=$?"
ls -la "$OUT" | tail -20
