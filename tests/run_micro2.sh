#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -x -q -k "roi or fast" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 python tests/bench_micro.py roialign > gpurun_out/micro.jsonl 2> gpurun_out/micro.err; echo "micro exit=$?"; cut -c1-330 gpurun_out/micro.jsonl; tail -5 gpurun_out/micro.err
timeout 400 python tests/gpu_engine_probe.py timing 8 > gpurun_out/timing8.log 2>&1; grep -E "roi|TOTAL" gpurun_out/timing8.log
