mkdir -p gpurun_out/r2f
for t in 128_4 128_6 128_8 256_2 256_3 256_4 512_1 512_2; do
  for c in "2 1 11115556 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
    cfg=${c%% *}
    RB_VARIANT=tune_${cfg}_$t timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$t /" | cut -c1-150 >> gpurun_out/r2f/tune.log
  done
done
cat gpurun_out/r2f/tune.log
