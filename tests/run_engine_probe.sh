#!/usr/bin/env bash
mkdir -p gpurun_out
echo "##### parity" > gpurun_out/engine.log
timeout 600 python tests/gpu_engine_probe.py parity >> gpurun_out/engine.log 2>&1
echo "exit=$?" >> gpurun_out/engine.log
echo "##### conv re-check (split accumulators)" >> gpurun_out/engine.log
timeout 300 python tests/gpu_probe.py conv_spatial >> gpurun_out/engine.log 2>&1
for b in 1 8; do
  echo "##### timing batch $b" >> gpurun_out/engine.log
  timeout 600 python tests/gpu_engine_probe.py timing $b >> gpurun_out/engine.log 2>&1
  echo "exit=$?" >> gpurun_out/engine.log
done
echo "##### timing batch 8 single-pass tf32" >> gpurun_out/engine.log
timeout 600 python tests/gpu_engine_probe.py timing 8 1 >> gpurun_out/engine.log 2>&1
tail -n 150 gpurun_out/engine.log
