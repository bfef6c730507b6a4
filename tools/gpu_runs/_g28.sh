mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/g28_bench_4gpu.json 2> gpurun_out/g28_err.txt; echo "rc=$?"
python - <<'PY'
import json
txt=open('gpurun_out/g28_bench_4gpu.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','replicas_identical','clocks')}, d['e2e']['value'])
PY
tail -n 3 gpurun_out/g28_err.txt | cut -c1-200
