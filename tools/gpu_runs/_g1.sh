mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/g1_smi.txt 2>&1
for args in "--d 128 --g 4 --bs 16 --S 300 --N 9" "--d 64 --g 4 --bs 16 --S 300 --N 9" "--d 128 --g 4 --bs 16 --S 1100 --N 40 --B 1"; do
  echo "== $args" >> gpurun_out/g1_attn_check.txt
  timeout 120 python tools/dflash_attn_check.py $args >> gpurun_out/g1_attn_check.txt 2>&1; echo "rc=$?" >> gpurun_out/g1_attn_check.txt
done
SF_DFLASH_TC_TESTS=1 timeout 300 python -m pytest tests/test_dflash_gpu.py -q -k tc_shapes > gpurun_out/g1_tc_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/g1_tc_tests.txt
tail -5 gpurun_out/g1_attn_check.txt; tail -5 gpurun_out/g1_tc_tests.txt
