mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g8_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g8_pytest.txt
tail -n 4 gpurun_out/g8_pytest.txt
timeout 900 python bench.py --steps 8 --warmup 3 --ab gemm_wide=0,1 --ab no_rope_fusion=0,1 --ab-rounds 6 --timeline gpurun_out/g8_timeline.json > gpurun_out/g8_bench.txt 2> gpurun_out/g8_bench_ab.txt; echo "rc=$?" >> gpurun_out/g8_bench.txt
tail -n 3 gpurun_out/g8_bench_ab.txt; cut -c1-250 gpurun_out/g8_bench.txt
timeout 600 python bench.py --workload dflash --steps 3 --warmup 3 > gpurun_out/g8_bench_dflash.txt 2>&1; echo "rc=$?" >> gpurun_out/g8_bench_dflash.txt
cut -c1-250 gpurun_out/g8_bench_dflash.txt
