#!/bin/bash
# One GPU-box visit of a round: parity tests (+ their measured-value logs), the bench line with every extra workload, the reference
# arm, the ncu launch list and `ncu --set full` captures of the step's dominant kernels and of the HBM-bound small kernels.  Run it
# THROUGH gpurun from the repo root, e.g.
#     gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02'            # everything (about 8 GPU-minutes)
#     gpurun --timeout 600 -- 'bash tools/gpu_round.sh r02 quick'       # tests + default bench + launch list
# Everything lands in gpurun_out/<round>_*; copy what should be judged into profiles/ (python tools/launch_table.py,
# python tools/ncu_summary.py <rep> print the summaries).  Numbers printed by a run under ncu are never bench values.
set -u
R=${1:-r02}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p "$OUT"

timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/${R}_gpu_tests.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/${R}_gpu_tests.log"
tail -3 "$OUT/${R}_gpu_tests.log"
# the measured values behind the tolerances (printed by the tests under -s): teacher-forced / free-running 100 steps, full sizes
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py -m gpu -q -s -p no:cacheprovider \
    -k "100_steps or full_size or variants_teacher_forced" 2>&1 | grep -oE "\[(teacher-forced|free-running|full-size)[^]]*\].*|[0-9]+ (passed|failed).*" > "$OUT/${R}_parity_values.log"
echo "parity values rc=$?"; tail -4 "$OUT/${R}_parity_values.log"

timeout 600 python bench.py > "$OUT/${R}_bench_g4096.json" 2> "$OUT/${R}_bench_g4096.err"
echo "bench rc=$?"; cut -c1-400 "$OUT/${R}_bench_g4096.json"

# per-launch durations of one eager step (same command as the bench, two timed steps, no graph so every kernel is listed)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/${R}_launches_g4096.csv" \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > "$OUT/${R}_ncu_launches.log" 2>&1
echo "launch list rc=$?"

[ "$MODE" = quick ] && exit 0

timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/${R}_bench_reference.json" 2> "$OUT/${R}_bench_reference.err"
echo "reference arm rc=$?"

# full captures.  Launch arithmetic of `bench.py --no-graph` (30 CE-only pretraining steps = 1 forward + 1 backward each, then 3
# warm-up steps of 2 forwards (clean, fused noisy) + 2 backwards): the first TIMED step's four encoder launches are launches 72..75
# of the encoder kernels; its three ECDA tensor-core / elementwise kernels come after the 9 of the warm-up steps.
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_enc_(fwd|bwd)_t[cp]" --launch-skip 72 -c 4 \
    -o "$OUT/${R}_full_encoders" -f python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > "$OUT/${R}_ncu_full_encoders.log" 2>&1
echo "ncu full (encoders) rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_gram_tc|k_pz_tc|k_ecda_p" --launch-skip 9 -c 3 \
    -o "$OUT/${R}_full_ecda" -f python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > "$OUT/${R}_ncu_full_ecda.log" 2>&1
echo "ncu full (ECDA) rc=$?"
# HBM-bound small kernels at the headline size (achieved DRAM GB/s): plain run first, then the capture of the same command
timeout 300 python tools/prof_small.py 4096 200 > "$OUT/${R}_prof_small_plain.log" 2>&1 &&
timeout 500 ncu --set full --clock-control none -k "regex:k_augment|k_gather_pad|k_pool_finalize|k_grad_reduce|k_adam_ema|k_bwd_db" --launch-skip 8 -c 14 \
    -o "$OUT/${R}_full_small" -f python tools/prof_small.py 4096 200 > "$OUT/${R}_ncu_full_small.log" 2>&1
echo "ncu full (small kernels) rc=$?"
ls -la "$OUT" | grep "${R}_" | tail -20
