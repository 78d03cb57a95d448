#!/usr/bin/env bash
# gpurun --gpus N helper: SURVEY 8e on hardware (NCCL gather == single-GPU run, bit-exact) and BASELINE configs[3]
# (Mask R-CNN R-101-FPN, 8 images per GPU) at N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "multi_gpu_gather" > gpurun_out/r02_pytest_dist_n$N.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/r02_pytest_dist_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 tests/dist_gather_check.py > gpurun_out/r02_dist_gather_n$N.json 2> gpurun_out/r02_dist_n$N.err; echo "gather exit=$?"; tail -n 1 gpurun_out/r02_dist_gather_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus $N --steps 40 --warmup 3 --arch resnet101 --no-microbench > gpurun_out/r02_bench_r101_n$N.json 2>> gpurun_out/r02_dist_n$N.err; echo "r101 exit=$?"; cut -c1-400 gpurun_out/r02_bench_r101_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus $N --steps 40 --warmup 3 --no-microbench > gpurun_out/r02_bench_n$N.json 2>> gpurun_out/r02_dist_n$N.err; echo "r50 exit=$?"; cut -c1-400 gpurun_out/r02_bench_n$N.json
tail -n 5 gpurun_out/r02_dist_n$N.err
