mkdir -p gpurun_out
for i in 1 2; do
for cfg in "8 0" "16 0" "8 8" "16 8" "16 16"; do
set -- $cfg
DT_CONV_RING_KB=$1 DT_CONV_RING2_KB=$2 timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_rs_$1_$2_$i.log 2>&1; echo "ring_kb=$1 ring2_kb=$2 $(tail -n 2 gpurun_out/ops_rs_$1_$2_$i.log | head -1)"
done
done
