mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -q -x -k "gemm or golden" > gpurun_out/g10_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g10_pytest.txt
tail -n 4 gpurun_out/g10_pytest.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g10_bench.txt 2> gpurun_out/g10_bench_err.txt; echo "rc=$?" >> gpurun_out/g10_bench.txt
cut -c1-250 gpurun_out/g10_bench.txt
