#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tests/gpu_engine_probe.py parity > gpurun_out/parity_fpn.txt 2>&1; tail -n 6 gpurun_out/parity_fpn.txt
bash tests/run_round.sh ncu
