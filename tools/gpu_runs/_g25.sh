mkdir -p gpurun_out
timeout 600 python bench.py --workload dflash --steps 5 --warmup 3 --no-cpu-baseline --timeline gpurun_out/g25_dflash_timeline.json > gpurun_out/g25_bench_dflash.json 2> gpurun_out/g25_err.txt; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g25_bench_dflash.json').readline())
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['step_frac_of_burst'], d['roofline']['gemm_ms_per_step'])
t=json.load(open('gpurun_out/g25_dflash_timeline.json'))
for x in t['by_kernel_ms'][:24]: print(round(x['ms']/2,2), x['n']//2, x['name'][:80])
print(sum(x['ms'] for x in t['by_kernel_ms'])/2, t['idle_ms'])
PY
