mkdir -p gpurun_out/r2t
for t in 256_4_4 256_4_3 256_4_2 128_8_4 128_8_2; do
  for c in "2 1 11115556 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
    cfg=${c%% *}
    RB_VARIANT=tune_${cfg}_$t timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$t /" | cut -c1-170 >> gpurun_out/r2t/tune.log
  done
done
cat gpurun_out/r2t/tune.log
