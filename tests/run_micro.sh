#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 python tests/bench_micro.py > gpurun_out/micro.jsonl 2> gpurun_out/micro.err; echo "micro exit=$?"; cat gpurun_out/micro.jsonl; tail -5 gpurun_out/micro.err
timeout 400 python tests/gpu_engine_probe.py timing 8 > gpurun_out/timing8.log 2>&1; tail -n 16 gpurun_out/timing8.log
