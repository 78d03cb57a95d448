mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "conv_kernel_variants and (2sm or RPN or default)" 2>&1 | tail -n 3
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_final.json'));print(d['value'],d['e2e']['value'],d['gpu_launches'],d['roofline']['frac'],d['roofline']['frac_executed'])"
