mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -q -x > gpurun_out/g19_pytest_dp.txt 2>&1; echo "rc=$?" >> gpurun_out/g19_pytest_dp.txt
tail -n 5 gpurun_out/g19_pytest_dp.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/g19_bench_2gpu.json 2> gpurun_out/g19_bench_2gpu_err.txt; echo "rc=$?"
cut -c1-400 gpurun_out/g19_bench_2gpu.json; tail -n 5 gpurun_out/g19_bench_2gpu_err.txt | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g19_bench_2gpu.json').readline())
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','replicas_identical','e2e','clocks')})
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/g19_smoke.txt 2>&1; tail -n 3 gpurun_out/g19_smoke.txt
