mkdir -p gpurun_out
rm -f gpurun_out/logits_parity.jsonl
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "wide" > gpurun_out/g4_pytest_wide.txt 2>&1; echo "rc=$?" >> gpurun_out/g4_pytest_wide.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g4_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g4_pytest.txt
cp gpurun_out/logits_parity.jsonl gpurun_out/g4_logits_parity.jsonl
timeout 500 python tools/gemm_vs_cublas.py --rounds 2 --variants gemm_wide=2 --out gpurun_out/g4_gemm_vs_cublas.json > gpurun_out/g4_gemm_vs_cublas.txt 2>&1; echo "rc=$?" >> gpurun_out/g4_gemm_vs_cublas.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ab gemm_wide=0,1 --ab-rounds 6 > gpurun_out/g4_bench.txt 2> gpurun_out/g4_bench_ab.txt; echo "rc=$?" >> gpurun_out/g4_bench.txt
tail -n 5 gpurun_out/g4_pytest_wide.txt; tail -n 25 gpurun_out/g4_pytest.txt; tail -n 14 gpurun_out/g4_gemm_vs_cublas.txt; tail -n 3 gpurun_out/g4_bench_ab.txt; cut -c1-200 gpurun_out/g4_bench.txt
