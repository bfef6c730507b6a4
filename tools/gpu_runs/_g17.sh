mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/g17_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g17_pytest.txt
tail -n 5 gpurun_out/g17_pytest.txt
timeout 900 python bench.py > gpurun_out/g17_bench.json 2> gpurun_out/g17_bench_err.txt; echo "rc=$?"
cut -c1-300 gpurun_out/g17_bench.json
timeout 600 python bench.py --workload dflash --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g17_bench_dflash.json 2>> gpurun_out/g17_bench_err.txt
cut -c1-200 gpurun_out/g17_bench_dflash.json
timeout 600 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/g17_bench_cfg3.json 2>> gpurun_out/g17_bench_err.txt
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/g17_bench_cfg5.json 2>> gpurun_out/g17_bench_err.txt
cut -c1-200 gpurun_out/g17_bench_cfg3.json; cut -c1-200 gpurun_out/g17_bench_cfg5.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/g17_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g17_ncu_launches_out.txt 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_kernel|rmsnorm|attn_|diag_scores|adamw|teacher_merge|rope_kernel|sqnorm|cvt_' -s 100 -c 100 -o gpurun_out/g17_step_nongemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g17_ncu_nongemm_out.txt 2>&1; echo "ncu nongemm rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'gemm_' -s 78 -c 78 -o gpurun_out/g17_step_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g17_ncu_gemm_out.txt 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'gemm_|nvjet' -o gpurun_out/g17_gemm_cmp -f python tools/gemm_ncu_compare.py > gpurun_out/g17_ncu_cmp_out.txt 2>&1; echo "ncu cmp rc=$?"
ls -la gpurun_out/*.ncu-rep
