#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k roi_align > gpurun_out/pytest_roi.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_roi.log
for nw in 16 24 8; do
DT_ROI_SMEM_WARPS=$nw timeout 600 python tests/bench_micro.py roialign > gpurun_out/micro_roi_w$nw.jsonl 2> gpurun_out/micro_roi.err; echo "warps $nw"; grep -o '"impl": "[a-z_]*"\|"pooled": [0-9]*\|"ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/micro_roi_w$nw.jsonl | paste - - - - | grep fast; tail -3 gpurun_out/micro_roi.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_smem_map --launch-skip 3 --launch-count 1 -o gpurun_out/prof_roi_smem2 python tests/bench_micro.py roialign > gpurun_out/ncu_roi.log 2>&1; echo "ncu smem exit=$?"
