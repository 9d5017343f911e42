mkdir -p gpurun_out/r3l
for e in "RB_FUSED_BOUNCE=1" "RB_X=0"; do
for c in "1 0 9000000 3" "1 0 1000000 5" "3 0 25000000 3" "2 1 11115556 3" "4 0 10000000 3"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-150 >> gpurun_out/r3l/survey.log
done
done
cat gpurun_out/r3l/survey.log
