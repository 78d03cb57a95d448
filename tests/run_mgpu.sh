#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k nms > gpurun_out/pytest_nms.log 2>&1; echo "pytest nms exit=$?"; tail -n 5 gpurun_out/pytest_nms.log
timeout 600 python tests/bench_micro.py nms > gpurun_out/micro_nms.jsonl 2>&1; cat gpurun_out/micro_nms.jsonl | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit=$?"; cat gpurun_out/bench_n2.json | cut -c1-700; tail -n 8 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2>> gpurun_out/bench_n2.err; echo "ref n2 exit=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json
