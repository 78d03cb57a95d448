# gpurun helper (1 GPU): one --set full capture (with source-level stall sampling) of the first conv launches of a step: stem, conv1, conv2, downsample, conv3
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-microbench --no-reference-flow"
DT_NCU_REGION=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel --launch-count ${1:-5} -o gpurun_out/r02_prof_first $B --steps 1 --warmup 3 > gpurun_out/ncu_first.log 2>&1
echo "ncu exit=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
