#!/bin/bash
# Runs each gemm_check group in its own process so one trapped kernel does not hide the others.
mkdir -p gpurun_out
for g in gemm conv epilogue batched wgrad wgrad_batched; do
  timeout 300 python tools/gemm_check.py $g 2>&1 | tail -40
done | tee gpurun_out/gemm_check.log
