# gpurun helper: full GPU test suite + smoke + the default bench line (no ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r02_pytest_gpu_final.log
tail -n 3 gpurun_out/r02_pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit=$?"; cut -c1-300 gpurun_out/r02_bench_final.json; tail -n 2 gpurun_out/r02_bench_final.err
