#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 600 python tests/gpu_engine_probe.py parity > gpurun_out/parity_fpn.txt 2>&1; grep "CMP\|FAIL\|Error" gpurun_out/parity_fpn.txt | tail -30
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
DT_CONV_KIND=tf32 timeout 900 python bench.py --no-cpu-baseline --steps 40 > gpurun_out/bench_tf32.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_tf32.json | cut -c1-300
timeout 600 python tests/gpu_engine_probe.py ops > gpurun_out/ops.log 2>&1; tail -n 3 gpurun_out/ops.log
