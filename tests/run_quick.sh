#!/usr/bin/env bash
mkdir -p gpurun_out
: > gpurun_out/probe.log
for g in conv_basic conv_spatial conv_epilogue; do
  echo "##### $g" >> gpurun_out/probe.log
  timeout 150 python tests/gpu_probe.py $g >> gpurun_out/probe.log 2>&1
  echo "exit=$?" >> gpurun_out/probe.log
done
grep -E "CONV|exit|bad idx|Error|error" gpurun_out/probe.log | cut -c1-190
if grep -q "exit=124" gpurun_out/probe.log; then echo "HANG detected, stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log
timeout 400 python tests/gpu_engine_probe.py ops > gpurun_out/ops.log 2>&1; tail -n 100 gpurun_out/ops.log
