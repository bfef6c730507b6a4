mkdir -p gpurun_out
rm -f gpurun_out/logits_parity.jsonl
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" > gpurun_out/g5_pytest_gemm.txt 2>&1; echo "rc=$?" >> gpurun_out/g5_pytest_gemm.txt
tail -n 5 gpurun_out/g5_pytest_gemm.txt
timeout 600 python tools/gemm_vs_cublas.py --rounds 2 --variants gemm_wide=2 gemm_epi_staged=1 gemm_wide=-1 --out gpurun_out/g5_gemm_vs_cublas.json > gpurun_out/g5_gemm_vs_cublas.txt 2>&1; echo "rc=$?" >> gpurun_out/g5_gemm_vs_cublas.txt
tail -n 14 gpurun_out/g5_gemm_vs_cublas.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ab gemm_wide=-1,0,2 --ab gemm_epi_staged=0,1 --ab-rounds 5 > gpurun_out/g5_bench.txt 2> gpurun_out/g5_bench_ab.txt; echo "rc=$?" >> gpurun_out/g5_bench.txt
tail -n 4 gpurun_out/g5_bench_ab.txt; cut -c1-200 gpurun_out/g5_bench.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g5_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g5_pytest.txt
tail -n 12 gpurun_out/g5_pytest.txt
