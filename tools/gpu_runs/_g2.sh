mkdir -p gpurun_out
timeout 300 python tools/gemm_vs_cublas.py --out gpurun_out/g2_gemm_vs_cublas.json > gpurun_out/g2_gemm_vs_cublas.txt 2>&1; echo "rc=$?" >> gpurun_out/g2_gemm_vs_cublas.txt
timeout 200 python tools/dflash_bench.py > gpurun_out/g2_dflash_bench_tc.txt 2>&1; echo "rc=$?" >> gpurun_out/g2_dflash_bench_tc.txt
SF_DFLASH_ATTN_TC=-1 timeout 200 python tools/dflash_bench.py > gpurun_out/g2_dflash_bench_cc.txt 2>&1; echo "rc=$?" >> gpurun_out/g2_dflash_bench_cc.txt
timeout 400 python bench.py --workload dflash --steps 3 --warmup 3 > gpurun_out/g2_bench_dflash.txt 2>&1; echo "rc=$?" >> gpurun_out/g2_bench_dflash.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g2_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g2_pytest.txt
tail -3 gpurun_out/g2_dflash_bench_tc.txt gpurun_out/g2_bench_dflash.txt gpurun_out/g2_pytest.txt
