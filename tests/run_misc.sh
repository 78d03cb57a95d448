#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tests/gpu_engine_probe.py c4 > gpurun_out/c4.log 2>&1; grep C4 gpurun_out/c4.log; tail -n 3 gpurun_out/c4.log
timeout 900 python bench.py --arch resnet101 --steps 40 --no-cpu-baseline > gpurun_out/bench_r101.json 2> gpurun_out/bench_r101.err; cut -c1-400 gpurun_out/bench_r101.json; tail -n 3 gpurun_out/bench_r101.err
timeout 900 python tests/bench_micro.py roialign > gpurun_out/micro.jsonl 2> gpurun_out/micro.err; cut -c1-300 gpurun_out/micro.jsonl
