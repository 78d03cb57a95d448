#!/usr/bin/env bash
# gpurun helper: GPU test suite, bench, ncu launch list, full ncu captures of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 python tests/gpu_engine_probe.py ops > gpurun_out/ops.log 2>&1; tail -n 3 gpurun_out/ops.log
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 240 --launch-count 290 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches exit=$?"; wc -l gpurun_out/launches.csv
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel --launch-skip 300 --launch-count 4 -o gpurun_out/prof_conv python tests/gpu_engine_probe.py timing 8 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?"; ls -la gpurun_out/
fi
