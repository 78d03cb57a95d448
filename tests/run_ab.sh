# gpurun helper: parity tests, then a same-box interleaved per-op A/B of one environment switch:  bash tests/run_ab.sh VAR OFF ON [tag]
VAR=$1; OFF=$2; ON=$3; TAG=${4:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit=$?"; tail -n 6 gpurun_out/${TAG}_pytest.log
for i in 1 2; do
env $VAR=$OFF timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_${TAG}_off_$i.log 2>&1; tail -n 1 gpurun_out/ops_${TAG}_off_$i.log
env $VAR=$ON timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_${TAG}_on_$i.log 2>&1; tail -n 1 gpurun_out/ops_${TAG}_on_$i.log
done
