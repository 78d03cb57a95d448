#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "teacher_forced_proposals or end_to_end" 2>&1 | tail -3
for kb in 56 28 14 7; do
  echo "== tile_kb $kb"; DT_ROI_TILE_KB=$kb timeout 300 python tests/bench_micro.py roialign 2>/dev/null | grep '"fast"' | cut -c1-260
done
timeout 300 python tests/gpu_engine_probe.py timing 8 2>&1 | grep -E "proposals|TOTAL"
