mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dflash_gpu.py -q -x > gpurun_out/g12_pytest_dflash.txt 2>&1; echo "rc=$?" >> gpurun_out/g12_pytest_dflash.txt
tail -n 6 gpurun_out/g12_pytest_dflash.txt
for v in 0 2 1; do
SF_GEMM_EPI_STAGED=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/g12_bench_staged$v.txt 2>/dev/null; echo "rc=$?" >> gpurun_out/g12_bench_staged$v.txt
done
python - <<'PY'
import json
for v in (0,2,1):
    ls=[x for x in open(f'gpurun_out/g12_bench_staged{v}.txt') if x.startswith('{')]
    d=json.loads(ls[0]); print('staged',v, round(d['ms_per_step'],1), d['clocks']['sm_mhz'], [(s['N'],s['K'],s['tflops']) for s in d['roofline']['gemm_by_shape'] if s['K']==4096])
PY
