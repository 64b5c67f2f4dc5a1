#!/usr/bin/env bash
# Runs each diagnostic section in its own process with its own timeout, so one hang cannot mask the rest.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for sec in gemm forward_small_bf16 loop_small_bf16 forward_small_fp16 loop_small_fp16 c1 timing; do
  echo "=== $sec"
  timeout ${DIAG_TIMEOUT:-240} python tools/gpu_diag.py $sec > gpurun_out/diag_$sec.log 2>&1
  echo "exit=$?"
  tail -c 2500 gpurun_out/diag_$sec.log
done
