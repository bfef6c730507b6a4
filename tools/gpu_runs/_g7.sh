mkdir -p gpurun_out
timeout 120 python tools/gemm_trace.py 16384 4096 4096 > gpurun_out/g7_trace_o.txt 2>&1; echo "rc=$?" >> gpurun_out/g7_trace_o.txt
timeout 120 python tools/gemm_trace.py 16384 4096 12288 > gpurun_out/g7_trace_down.txt 2>&1
cat gpurun_out/g7_trace_o.txt | head -16
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" > gpurun_out/g7_pytest_gemm.txt 2>&1; echo "rc=$?" >> gpurun_out/g7_pytest_gemm.txt
tail -n 4 gpurun_out/g7_pytest_gemm.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ab gemm_wide=0,1,-1 --ab no_rope_fusion=0,1 --ab-rounds 8 > gpurun_out/g7_bench.txt 2> gpurun_out/g7_bench_ab.txt; echo "rc=$?" >> gpurun_out/g7_bench.txt
tail -n 3 gpurun_out/g7_bench_ab.txt; cut -c1-200 gpurun_out/g7_bench.txt
