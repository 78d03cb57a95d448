#!/usr/bin/env bash
mkdir -p gpurun_out
for m in auto 1sm 2sm; do
  if [ $m == auto ]; then unset DT_CONV_MMA; else export DT_CONV_MMA=$m; fi
  timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_$m.log 2>&1; tail -n 2 gpurun_out/ops_$m.log
done
