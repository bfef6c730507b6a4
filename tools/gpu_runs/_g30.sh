mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/g30_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/g30_pytest.txt
tail -n 4 gpurun_out/g30_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
