mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py -m gpu -x -q > gpurun_out/r2g/pytest.log 2>&1
tail -3 gpurun_out/r2g/pytest.log
for t in default 128_4 128_6 128_8 256_2 256_3 256_4 512_1 512_2; do
  for c in "1 0 9000000 3" "2 1 11115556 3" "3 0 9000000 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
    cfg=${c%% *}
    if [ $t = default ]; then timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$t /" | cut -c1-170 >> gpurun_out/r2g/tune.log
    elif [ $cfg = 2 -o $cfg = 4 -o $cfg = 5 ]; then RB_VARIANT=tune_${cfg}_$t timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$t /" | cut -c1-170 >> gpurun_out/r2g/tune.log; fi
  done
done
cat gpurun_out/r2g/tune.log
