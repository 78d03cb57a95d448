#!/usr/bin/env bash
mkdir -p gpurun_out
: > gpurun_out/probe.log
for g in conv_basic conv_spatial conv_epilogue; do
  echo "##### $g (2sm)" >> gpurun_out/probe.log
  timeout 120 python tests/gpu_probe.py $g >> gpurun_out/probe.log 2>&1
  echo "exit=$?" >> gpurun_out/probe.log
done
grep -E "CONV|exit|bad idx|rror" gpurun_out/probe.log | cut -c1-200
if grep -q "exit=124" gpurun_out/probe.log; then echo "HANG detected in 2sm mode"; fi
if grep -E "exit=(124|1)$" gpurun_out/probe.log > /dev/null; then
  echo "2sm mode failing; 1sm sanity:"; DT_CONV_1SM=1 timeout 120 python tests/gpu_probe.py conv_basic 2>&1 | grep -E "CONV|rror" | cut -c1-160
  exit 0
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 400 python tests/gpu_engine_probe.py ops > gpurun_out/ops_2sm.log 2>&1; tail -n 2 gpurun_out/ops_2sm.log
DT_CONV_1SM=1 timeout 400 python tests/gpu_engine_probe.py ops > gpurun_out/ops_1sm.log 2>&1; tail -n 2 gpurun_out/ops_1sm.log
