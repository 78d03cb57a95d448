#!/usr/bin/env bash
# gpurun helper: each probe group under its own timeout; logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "##### $g" >> gpurun_out/probe.log
  timeout 240 python tests/gpu_probe.py $g >> gpurun_out/probe.log 2>&1
  echo "exit=$?" >> gpurun_out/probe.log
done
tail -n 120 gpurun_out/probe.log
