mkdir -p gpurun_out
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --ab gemm_epi_staged=0,1,2 --ab gemm_epi8=0,-1 --ab-rounds 10 > gpurun_out/g13_bench.txt 2> gpurun_out/g13_bench_ab.txt; echo "rc=$?" >> gpurun_out/g13_bench.txt
grep '"ab"' gpurun_out/g13_bench_ab.txt; cut -c1-200 gpurun_out/g13_bench.txt
SF_GEMM_EPI_STAGED=1 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --ab gemm_epi8=0,-1 --ab-rounds 10 > gpurun_out/g13_bench_s1.txt 2> gpurun_out/g13_bench_s1_ab.txt
grep '"ab"' gpurun_out/g13_bench_s1_ab.txt; cut -c1-200 gpurun_out/g13_bench_s1.txt
timeout 900 python tools/reference_gpu_bench.py --backend flex_attention sdpa --batch 8 --steps 2 --warmup 1 > gpurun_out/g13_reference_gpu.txt 2> gpurun_out/g13_reference_gpu_err.txt; echo "rc=$?" >> gpurun_out/g13_reference_gpu.txt
cut -c1-600 gpurun_out/g13_reference_gpu.txt; tail -n 5 gpurun_out/g13_reference_gpu_err.txt | cut -c1-300
