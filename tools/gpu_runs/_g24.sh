mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-variants > gpurun_out/g24_bench.json 2> gpurun_out/g24_bench_err.txt; echo "rc=$?"
grep e2e_variants gpurun_out/g24_bench_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g24_bench.json').readline())
print(d['ms_per_step'], d['e2e'], d['clocks'])
PY
