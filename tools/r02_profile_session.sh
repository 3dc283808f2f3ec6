#!/bin/bash
# One GPU session of round 2: tests, ncu traffic / launch captures of one train_step, compute-sanitizer on tiny shapes.
# Everything lands in gpurun_out/; summaries are made afterwards with tools/summarize_ncu.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02g_tests.log 2>&1; echo "exit $?" >> gpurun_out/r02g_tests.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
for dt in float32 bfloat16; do
  timeout 600 ncu --metrics $M --clock-control none -k "regex:gemm_fwd_kernel|conv3x3_resident_kernel" -c 270 --csv \
    --log-file gpurun_out/r02_traffic_$dt.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline \
    --resnet-dtype $dt > gpurun_out/r02_traffic_$dt.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline \
  > gpurun_out/r02_launches.log 2>&1
# racecheck / memcheck of the hand-rolled mbarrier / TMEM pipelines at tiny shapes
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "test_conv_forward_and_wgrad_match_oracle or test_subpixel_conv_equals_upsample_then_conv" \
  > gpurun_out/r02_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r02_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "test_train_step_matches_oracle_for_two_steps or test_conv_forward_and_wgrad_match_oracle" \
  > gpurun_out/r02_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r02_memcheck.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
grep -n "passed\|failed\|^FAILED" gpurun_out/r02g_tests.log | tail -n 20
tail -n 5 gpurun_out/r02_racecheck.log gpurun_out/r02_memcheck.log
cut -c1-600 gpurun_out/r02g_bench.json
