mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/halo_pytest.log 2>&1; echo "pytest exit=$?"; tail -n 15 gpurun_out/halo_pytest.log
for i in 1 2; do
DT_CONV_HALO=0 timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_halo_off_$i.log 2>&1; tail -n 1 gpurun_out/ops_halo_off_$i.log
DT_CONV_HALO=1 timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_halo_on_$i.log 2>&1; tail -n 1 gpurun_out/ops_halo_on_$i.log
done
