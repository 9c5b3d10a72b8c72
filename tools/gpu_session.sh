#!/bin/bash
# One GPU-box session: run every diagnostic group in its own process under a timeout; logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/gpu.txt
for g in "$@"; do
  echo "##### $g"
  timeout 300 python tools/gpu_check.py $g 2>&1 | tee gpurun_out/check_$g.log | tail -40
done
