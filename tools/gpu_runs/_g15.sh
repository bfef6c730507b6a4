mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/g15_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g15_pytest.txt
tail -n 5 gpurun_out/g15_pytest.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --timeline gpurun_out/g15_timeline.json > gpurun_out/g15_bench.txt 2> gpurun_out/g15_bench_err.txt; echo "rc=$?" >> gpurun_out/g15_bench.txt
cut -c1-250 gpurun_out/g15_bench.txt
python - <<'PY'
import json
t=json.load(open('gpurun_out/g15_timeline.json'))
for x in t['by_kernel_ms'][:16]: print(round(x['ms']/2,2), x['n']//2, x['name'][:70])
PY
SF_GEMM_WIDE=2 timeout 120 python tools/gemm_trace.py 16384 4096 4096 > gpurun_out/g15_trace_o.txt 2>&1
head -12 gpurun_out/g15_trace_o.txt
