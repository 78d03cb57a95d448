#!/usr/bin/env bash
# gpurun helper: the round evidence run on 1 GPU (full GPU test suite, smoke, bench both arms, per-op events, ncu launch list with DRAM bytes,
# one --set full capture of conv launches and of the RoIAlign microbench kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r02_pytest_gpu_final.log
tail -n 4 gpurun_out/r02_pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit=$?"; cut -c1-400 gpurun_out/r02_bench_final.json; tail -n 3 gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>> gpurun_out/r02_bench_final.err; cut -c1-300 gpurun_out/r02_bench_ref_final.json
timeout 600 python bench.py --arch resnet101 --steps 40 --no-cpu-baseline --no-microbench --no-reference-flow > gpurun_out/r02_bench_r101_n1.json 2>> gpurun_out/r02_bench_final.err; cut -c1-200 gpurun_out/r02_bench_r101_n1.json
timeout 600 python tests/gpu_engine_probe.py ops > gpurun_out/r02_ops_final.log 2>&1; tail -n 2 gpurun_out/r02_ops_final.log
timeout 600 python tests/gpu_engine_probe.py c4 > gpurun_out/r02_c4_final.log 2>&1; tail -n 4 gpurun_out/r02_c4_final.log
B="python bench.py --no-cpu-baseline --no-microbench --no-reference-flow"
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_dram_final.csv $B --steps 1 --warmup 3 > gpurun_out/ncu_launches_final.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/r02_launches_dram_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_align_smem_map_kernel --launch-count 2 -o gpurun_out/r02_prof_roialign python tests/bench_micro.py roialign > gpurun_out/ncu_roi.log 2>&1
echo "ncu roialign exit=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel --launch-count 16 -o gpurun_out/r02_prof_conv $B --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu conv exit=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
