mkdir -p gpurun_out/r4b
for spl in 0 -1 2 3; do
for c in "1 0 9000000 3" "1 0 1000000 5" "3 0 25000000 3" "2 1 11115556 3"; do
  timeout 300 python profiles/trace_one.py $c spl=$spl 2>&1 | cut -c1-150 >> gpurun_out/r4b/survey.log
done
done
cat gpurun_out/r4b/survey.log
