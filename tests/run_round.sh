#!/usr/bin/env bash
# gpurun helper: the round evidence run on 1 GPU (tests, smoke, bench both arms, per-op events, C4, parity report, microbenchmarks, ncu launch list + full capture)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cut -c1-400 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --arch resnet101 --steps 40 --no-cpu-baseline > gpurun_out/bench_r101.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_r101.json
timeout 600 python tests/gpu_engine_probe.py ops > gpurun_out/ops.log 2>&1; tail -n 2 gpurun_out/ops.log
timeout 600 python tests/gpu_engine_probe.py c4 > gpurun_out/c4.log 2>&1; tail -n 6 gpurun_out/c4.log
timeout 600 python tests/gpu_engine_probe.py parity > gpurun_out/parity_fpn.txt 2>&1; grep "mask logits" gpurun_out/parity_fpn.txt
timeout 900 python tests/bench_micro.py > gpurun_out/micro.jsonl 2> gpurun_out/micro.err; echo "micro exit=$?"
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/launches.csv
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel --launch-skip 30 --launch-count 6 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; ls -la gpurun_out/ | head -40
