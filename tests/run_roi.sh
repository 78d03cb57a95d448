#!/usr/bin/env bash
# gpurun helper: RoIAlign tests + microbench (both fast variants) + one ncu capture of each
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k roi_align > gpurun_out/pytest_roi.log 2>&1; echo "pytest exit=$?"; tail -n 6 gpurun_out/pytest_roi.log
timeout 600 python tests/bench_micro.py roialign > gpurun_out/micro_roi.jsonl 2> gpurun_out/micro_roi.err; echo "micro exit=$?"; grep -o '"impl": "[a-z_]*"\|"pooled": [0-9]*\|"ms": [0-9.]*\|"frac": [0-9.]*\|"max_abs_diff_fast_vs_exact": [0-9.e-]*' gpurun_out/micro_roi.jsonl | paste - - - - - ; tail -3 gpurun_out/micro_roi.err
DT_ROI_SMEM_MAP=0 timeout 600 python tests/bench_micro.py roialign > gpurun_out/micro_roi_gather.jsonl 2>> gpurun_out/micro_roi.err; grep -o '"impl": "[a-z_]*"\|"pooled": [0-9]*\|"ms": [0-9.]*' gpurun_out/micro_roi_gather.jsonl | paste - - -
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_smem_map --launch-skip 3 --launch-count 1 -o gpurun_out/prof_roi_smem python tests/bench_micro.py roialign > gpurun_out/ncu_roi.log 2>&1; echo "ncu smem exit=$?"
  DT_ROI_SMEM_MAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_fast_nchw_out --launch-skip 3 --launch-count 1 -o gpurun_out/prof_roi_gather python tests/bench_micro.py roialign >> gpurun_out/ncu_roi.log 2>&1; echo "ncu gather exit=$?"
fi
