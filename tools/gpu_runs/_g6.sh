mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -q -x -k "gemm or golden" > gpurun_out/g6_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g6_pytest.txt
tail -n 6 gpurun_out/g6_pytest.txt
timeout 500 python tools/gemm_vs_cublas.py --rounds 2 --variants gemm_wide=-1 --out gpurun_out/g6_gemm_vs_cublas.json > gpurun_out/g6_gemm_vs_cublas.txt 2>&1; echo "rc=$?" >> gpurun_out/g6_gemm_vs_cublas.txt
tail -n 14 gpurun_out/g6_gemm_vs_cublas.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ab gemm_wide=-1,0,1 --ab no_teacher_fusion=0,1 --ab no_loss_stats_fusion=0,1 --ab-rounds 5 > gpurun_out/g6_bench.txt 2> gpurun_out/g6_bench_ab.txt; echo "rc=$?" >> gpurun_out/g6_bench.txt
tail -n 4 gpurun_out/g6_bench_ab.txt; cut -c1-200 gpurun_out/g6_bench.txt
timeout 600 ncu --set full --clock-control none -k regex:'gemm_|nvjet' -o gpurun_out/g6_gemm_cmp -f python tools/gemm_ncu_compare.py > gpurun_out/g6_ncu.txt 2>&1; echo "rc=$?" >> gpurun_out/g6_ncu.txt
tail -n 3 gpurun_out/g6_ncu.txt
