# gpurun helper: same-box interleaved per-op logs for several settings of one environment switch:  bash tests/run_ab_multi.sh TAG VAR v0 v1 v2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for i in 1 2; do
for v in "$@"; do
env $VAR=$v timeout 300 python tests/gpu_engine_probe.py ops > gpurun_out/ops_${TAG}_${v}_$i.log 2>&1; echo "$VAR=$v $(tail -n 2 gpurun_out/ops_${TAG}_${v}_$i.log | head -1)"
done
done
