mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dflash_gpu.py -q -x > gpurun_out/g16_pytest_dflash.txt 2>&1; echo "rc=$?" >> gpurun_out/g16_pytest_dflash.txt
tail -n 6 gpurun_out/g16_pytest_dflash.txt
( for w in 0 40 300 2048; do timeout 300 python tools/dflash_attn_check.py --S 1100 --N 24 --window $w; done
  timeout 300 python tools/dflash_attn_check.py --S 700 --N 16 --d 64 --window 130
  timeout 300 python tools/dflash_attn_check.py --S 2500 --N 40 --B 1 --window 512 ) > gpurun_out/g16_attn_check.txt 2>&1
grep -c "cos 0.99999\|cos 1.0000" gpurun_out/g16_attn_check.txt; grep -v "cos 0.99999\|cos 1.0000" gpurun_out/g16_attn_check.txt | head -20
timeout 600 python bench.py --workload dflash --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g16_bench_dflash.txt 2> gpurun_out/g16_bench_dflash_err.txt
cut -c1-300 gpurun_out/g16_bench_dflash.txt
