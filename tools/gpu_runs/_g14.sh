mkdir -p gpurun_out
timeout 120 python tools/gemm_trace.py 16384 4096 4096 > gpurun_out/g14_trace_o.txt 2>&1; echo "rc=$?" >> gpurun_out/g14_trace_o.txt
cat gpurun_out/g14_trace_o.txt | head -12
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" > gpurun_out/g14_pytest_gemm.txt 2>&1; echo "rc=$?" >> gpurun_out/g14_pytest_gemm.txt
tail -n 4 gpurun_out/g14_pytest_gemm.txt
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --ab gemm_wide=0,4 --ab-rounds 10 > gpurun_out/g14_bench.txt 2> gpurun_out/g14_bench_ab.txt; echo "rc=$?" >> gpurun_out/g14_bench.txt
grep '"ab"' gpurun_out/g14_bench_ab.txt; cut -c1-200 gpurun_out/g14_bench.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g14_bench.txt').readline())
print(d['ms_per_step'], d['clocks'], d['roofline'].get('gemm_by_shape'))
PY
timeout 600 python tools/gemm_vs_cublas.py --ms 400 --rounds 3 --variants gemm_wide=4 > gpurun_out/g14_gemm_vs_cublas.txt 2>&1
tail -n 30 gpurun_out/g14_gemm_vs_cublas.txt
