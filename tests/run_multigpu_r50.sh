#!/usr/bin/env bash
# gpurun --gpus N helper (short): the headline R-50-FPN bench at N GPUs without the CPU legs
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus $N --steps 40 --warmup 3 --no-microbench --no-cpu-baseline --no-reference-flow > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_dist_r50_n$N.err; echo "r50 exit=$?"; cut -c1-330 gpurun_out/r02_bench_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29674 bench.py --gpus $N --steps 40 --warmup 3 --arch resnet101 --no-microbench --no-cpu-baseline --no-reference-flow > gpurun_out/r02_bench_r101_n$N.json 2>> gpurun_out/r02_dist_r50_n$N.err; echo "r101 exit=$?"; cut -c1-330 gpurun_out/r02_bench_r101_n$N.json
