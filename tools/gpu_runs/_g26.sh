mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dflash_gpu.py -q -x > gpurun_out/g26_pytest_dflash.txt 2>&1; echo "rc=$?" >> gpurun_out/g26_pytest_dflash.txt
tail -n 4 gpurun_out/g26_pytest_dflash.txt
( timeout 300 python tools/dflash_attn_check.py --S 1100 --N 24 --window 0 --impl 1
  timeout 300 python tools/dflash_attn_check.py --S 1100 --N 24 --window 300 --impl 1
  timeout 300 python tools/dflash_attn_check.py --S 700 --N 16 --d 64 --window 130 --impl 1 ) > gpurun_out/g26_attn_check.txt 2>&1
grep -c "cos 0.99999\|cos 1.0000" gpurun_out/g26_attn_check.txt; grep -v "cos 0.99999\|cos 1.0000" gpurun_out/g26_attn_check.txt | head
timeout 600 python bench.py --workload dflash --steps 5 --warmup 3 --no-cpu-baseline --timeline gpurun_out/g26_dflash_timeline.json > gpurun_out/g26_bench_dflash.json 2> gpurun_out/g26_err.txt; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g26_bench_dflash.json').readline())
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['step_frac_of_burst'])
t=json.load(open('gpurun_out/g26_dflash_timeline.json'))
for x in t['by_kernel_ms']:
    if any(k in x['name'] for k in ('own','headnorm','ce_kernel','attn_fwd','ctx','bwd_dq')): print(round(x['ms']/2,2), x['n']//2, x['name'][:80])
print(sum(x['ms'] for x in t['by_kernel_ms'])/2)
PY
