mkdir -p gpurun_out
rm -f gpurun_out/logits_parity.jsonl
SF_NO_TEACHER_FUSION=1 SF_NO_LOSS_STATS_FUSION=1 timeout 1500 python -m pytest tests/test_step_gpu.py tests/test_ops_gpu.py -q -s -k "golden or step_shapes or swiglu_epilogues or ttt_attention" > gpurun_out/g3_pytest_unfused.txt 2>&1; echo "rc=$?" >> gpurun_out/g3_pytest_unfused.txt
mv gpurun_out/logits_parity.jsonl gpurun_out/g3_logits_parity_unfused.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g3_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g3_pytest.txt
timeout 600 ncu --set full --clock-control none -k regex:'gemm_kernel|nvjet|cutlass|sm100|xmma|cublas' -o gpurun_out/g3_gemm_cmp -f python tools/gemm_ncu_compare.py > gpurun_out/g3_ncu.txt 2>&1; echo "rc=$?" >> gpurun_out/g3_ncu.txt
tail -n 15 gpurun_out/g3_pytest_unfused.txt; tail -n 15 gpurun_out/g3_pytest.txt; tail -n 5 gpurun_out/g3_ncu.txt
timeout 400 python tools/gemm_vs_cublas.py --rounds 2 --variants gemm_stages=7 --out gpurun_out/g3_gemm_vs_cublas.json > gpurun_out/g3_gemm_vs_cublas.txt 2>&1; echo "rc=$?" >> gpurun_out/g3_gemm_vs_cublas.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ab gemm_stages=6,7 --ab-rounds 6 > gpurun_out/g3_bench.txt 2> gpurun_out/g3_bench_ab.txt; echo "rc=$?" >> gpurun_out/g3_bench.txt
tail -n 14 gpurun_out/g3_gemm_vs_cublas.txt; cat gpurun_out/g3_bench_ab.txt | tail -n 3; cut -c1-300 gpurun_out/g3_bench.txt
