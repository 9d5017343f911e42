mkdir -p gpurun_out/r3w
for v in "" "_256_4_s2" "_256_4_s3"; do
for c in "2 1 11115556 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  cfg=${c%% *}
  if [ -z "$v" ]; then e="RB_X=0"; else e="RB_VARIANT=tune_${cfg}${v}"; fi
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | cut -c1-170 >> gpurun_out/r3w/survey.log
done
done
cat gpurun_out/r3w/survey.log
