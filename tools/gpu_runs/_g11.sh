mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -q -x -k "gemm or golden" > gpurun_out/g11_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g11_pytest.txt
tail -n 4 gpurun_out/g11_pytest.txt
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --ab gemm_epi8=0,-1 --ab-rounds 8 > gpurun_out/g11_bench.txt 2> gpurun_out/g11_bench_ab.txt; echo "rc=$?" >> gpurun_out/g11_bench.txt
tail -n 2 gpurun_out/g11_bench_ab.txt; cut -c1-250 gpurun_out/g11_bench.txt
