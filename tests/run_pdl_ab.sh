# gpurun helper: PDL on/off A/B on the graph-replayed bench step (same box, interleaved)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-microbench --no-reference-flow --steps 60"
for i in 1 2; do
timeout 600 $B > gpurun_out/pdl0_$i.json 2> gpurun_out/pdl.err; python -c "import json;d=json.load(open('gpurun_out/pdl0_$i.json'));print('pdl0',d['ms_per_step'],d['e2e']['ms_per_step'])"
DT_CONV_PDL=1 timeout 600 $B > gpurun_out/pdl1_$i.json 2>> gpurun_out/pdl.err; python -c "import json;d=json.load(open('gpurun_out/pdl1_$i.json'));print('pdl1',d['ms_per_step'],d['e2e']['ms_per_step'])"
done
