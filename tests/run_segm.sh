#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/pytest_ops.log 2>&1; echo "pytest exit=$?"; tail -n 25 gpurun_out/pytest_ops.log
timeout 600 python tests/bench_micro.py roialign > gpurun_out/micro_roi.jsonl 2> gpurun_out/micro_roi.err; echo "micro exit=$?"; grep -o '"impl": "[a-z_]*"\|"pooled": [0-9]*\|"ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/micro_roi.jsonl | paste - - - - ; tail -3 gpurun_out/micro_roi.err
