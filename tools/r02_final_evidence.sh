#!/bin/bash
# Final evidence session of round 2: ncu launch list of one eager train_step (durations only) and compute-sanitizer
# racecheck / memcheck over the GEMM pipelines incl. the weight-gradient tap groups.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline \
  > gpurun_out/r02_final_launches.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py \
  tests/test_gpu_ops_r02.py -m gpu -q -x \
  -k "test_conv_forward_and_wgrad_match_oracle or test_subpixel_conv_equals_upsample_then_conv or test_wgrad_tap_groups_match_oracle" \
  > gpurun_out/r02_final_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r02_final_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py \
  tests/test_gpu_ops_r02.py -m gpu -q -x \
  -k "test_train_step_matches_oracle_for_two_steps or test_wgrad_tap_groups_match_oracle or test_pool_fused_conv_equals_conv_then_dsample" \
  > gpurun_out/r02_final_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r02_final_memcheck.log
tail -n 5 gpurun_out/r02_final_racecheck.log gpurun_out/r02_final_memcheck.log
wc -l gpurun_out/r02_final_launches.csv
