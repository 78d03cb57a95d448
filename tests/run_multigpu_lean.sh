#!/usr/bin/env bash
# gpurun --gpus N helper (short: N x box time is charged): SURVEY 8e on hardware (NCCL gather == single-GPU run, bit-exact) and
# BASELINE configs[3] (Mask R-CNN R-101-FPN, 8 images per GPU) at N GPUs, without the CPU legs of the bench
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 tests/dist_gather_check.py > gpurun_out/r02_dist_gather_n$N.json 2> gpurun_out/r02_dist_n$N.err; echo "gather exit=$?"; tail -n 1 gpurun_out/r02_dist_gather_n$N.json | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus $N --steps 40 --warmup 3 --arch resnet101 --no-microbench --no-cpu-baseline --no-reference-flow > gpurun_out/r02_bench_r101_n$N.json 2>> gpurun_out/r02_dist_n$N.err; echo "r101 exit=$?"; cut -c1-330 gpurun_out/r02_bench_r101_n$N.json
tail -n 3 gpurun_out/r02_dist_n$N.err
