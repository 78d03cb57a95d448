#!/usr/bin/env bash
# gpurun helper (1 GPU): PDL A/B on the graph-replayed step, ncu launch list + DRAM bytes of one step, one --set full capture of conv + RoIAlign launches
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-microbench --no-reference-flow"
timeout 600 $B --steps 60 > gpurun_out/r02_ab_pdl0.json 2> gpurun_out/r02_ab.err; cut -c1-260 gpurun_out/r02_ab_pdl0.json
DT_CONV_PDL=1 timeout 600 $B --steps 60 > gpurun_out/r02_ab_pdl1.json 2>> gpurun_out/r02_ab.err; echo "pdl exit=$?"; cut -c1-260 gpurun_out/r02_ab_pdl1.json
timeout 600 $B --steps 60 > gpurun_out/r02_ab_pdl0b.json 2>> gpurun_out/r02_ab.err; cut -c1-260 gpurun_out/r02_ab_pdl0b.json
DT_CONV_PDL=1 timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not notebook and not roialign_100k and not nms_100k" > gpurun_out/r02_pytest_pdl.log 2>&1; echo "pytest(pdl) exit=$?"; tail -n 3 gpurun_out/r02_pytest_pdl.log
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_dram.csv $B --steps 1 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/r02_launches_dram.csv
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel --launch-skip 3 --launch-count 12 -o gpurun_out/r02_prof_conv $B --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
