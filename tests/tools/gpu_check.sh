#!/bin/bash
# Runs the diagnostic stages in separate processes (a trap in one stage must not poison the next).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for s in simt umma fused net; do
  timeout 300 python tests/tools/gpu_check.py $s > gpurun_out/check_$s.log 2>&1
  echo "stage $s exit $?" | tee -a gpurun_out/check_summary.txt
  tail -40 gpurun_out/check_$s.log
done
