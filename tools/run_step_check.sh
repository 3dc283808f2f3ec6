#!/bin/bash
mkdir -p gpurun_out
for g in g d grads step; do
  timeout 600 python tools/step_check.py $g 2>&1 | tail -120
done | tee gpurun_out/step_check.log
