mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/g23_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g23_pytest.txt
tail -n 4 gpurun_out/g23_pytest.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --timeline gpurun_out/g23_timeline.json > gpurun_out/g23_bench.json 2> gpurun_out/g23_bench_err.txt; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g23_bench.json').readline())
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline'].get('step_frac_of_burst'), d['gpu_launches'], d['e2e']['value'])
t=json.load(open('gpurun_out/g23_timeline.json'))
for x in t['by_kernel_ms']:
    if any(k in x['name'] for k in ('teacher_merge','diag_scores','loss_kernel','rmsnorm','attn_fwd')): print(round(x['ms']/2,2), x['n']//2, x['name'][:70])
PY
