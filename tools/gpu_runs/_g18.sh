mkdir -p gpurun_out
export_rep() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; ls -la gpurun_out/$1.csv; }
timeout 900 python bench.py > gpurun_out/g18_bench.json 2> gpurun_out/g18_bench_err.txt; echo "bench rc=$?"
cut -c1-200 gpurun_out/g18_bench.json
timeout 600 python bench.py --workload dflash --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g18_bench_dflash.json 2>> gpurun_out/g18_bench_err.txt
timeout 600 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/g18_bench_cfg3.json 2>> gpurun_out/g18_bench_err.txt
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/g18_bench_cfg5.json 2>> gpurun_out/g18_bench_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/g18_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g18_ncu_launches_out.txt 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'loss_kernel|rmsnorm|attn_|diag_scores|rope_kernel|cvt_f32' -s 165 -c 36 -o gpurun_out/g18_step_nongemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g18_ncu_nongemm_out.txt 2>&1; echo "ncu nongemm rc=$?"
export_rep g18_step_nongemm
timeout 400 ncu --set full --clock-control none -k regex:'adamw|teacher_merge|sqnorm|cvt_flat|colsum|metrics_reduce|embedding|shift_left' -s 10 -c 10 -o gpurun_out/g18_step_small -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g18_ncu_small_out.txt 2>&1; echo "ncu small rc=$?"
export_rep g18_step_small
timeout 600 ncu --set full --clock-control none -k regex:'gemm_' -s 78 -c 20 -o gpurun_out/g18_step_gemm_fwd -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g18_ncu_gemm_out.txt 2>&1; echo "ncu gemm fwd rc=$?"
export_rep g18_step_gemm_fwd
timeout 600 ncu --set full --clock-control none -k regex:'gemm_' -s 138 -c 18 -o gpurun_out/g18_step_gemm_bwd -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline >> gpurun_out/g18_ncu_gemm_out.txt 2>&1; echo "ncu gemm bwd rc=$?"
export_rep g18_step_gemm_bwd
timeout 400 ncu --set full --clock-control none -k regex:'gemm_|nvjet' -o gpurun_out/g18_gemm_cmp -f python tools/gemm_ncu_compare.py > gpurun_out/g18_ncu_cmp_out.txt 2>&1; echo "ncu cmp rc=$?"
export_rep g18_gemm_cmp
du -sh gpurun_out
