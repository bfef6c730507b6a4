mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py tests/test_properties_gpu.py tests/test_dflash_gpu.py -q -x > gpurun_out/g29_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g29_pytest.txt
tail -n 4 gpurun_out/g29_pytest.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --ab rmsnorm_fwd_ring=0,-1 --ab-rounds 8 --timeline gpurun_out/g29_timeline.json > gpurun_out/g29_bench.json 2> gpurun_out/g29_err.txt; echo "rc=$?"
grep '"ab"' gpurun_out/g29_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g29_bench.json').readline())
print(d['value'], d['ms_per_step'], d['clocks'])
t=json.load(open('gpurun_out/g29_timeline.json'))
for x in t['by_kernel_ms']:
    if 'rmsnorm' in x['name']: print(round(x['ms']/2,2), x['n']//2, x['name'][:70])
PY
