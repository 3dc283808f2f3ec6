#!/bin/bash
# GPU session: full test suite, bench (default + bf16 ResNet), ncu capture of one step's forward-GEMM launches with
# DRAM bytes, duration and tensor-pipe activity (time-weighted tensor-pipe % over all conv launches).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02i}
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --resnet-dtype bfloat16 > gpurun_out/${TAG}_bench_bf16resnet.json 2>> gpurun_out/${TAG}_bench.err
# kernel durations: without programmatic dependent launch (a dependent kernel is resident, waiting, while its
# predecessor drains: CUPTI would count that wait as its duration)
XMC_PDL=0 python tools/gpu_idle.py 56 float32 > gpurun_out/${TAG}_insitu.md 2>&1
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -k "regex:gemm_fwd_kernel|conv3x3_resident_kernel|gemm_wgrad_kernel" -c 365 --csv \
  --log-file gpurun_out/${TAG}_convs.csv python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline \
  > gpurun_out/${TAG}_convs.log 2>&1
grep -n "passed\|failed\|^FAILED" gpurun_out/${TAG}_tests.log | tail -n 20
cut -c1-200 gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_bf16resnet.json
head -n 24 gpurun_out/${TAG}_insitu.md | tail -n 16
