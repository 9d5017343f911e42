mkdir -p gpurun_out/r2i
for c in "2 1 11115556 1" "4 0 10000000 1" "5 20 10000000 1 rings=10"; do
  cfg=${c%% *}
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2i/launches_cfg$cfg.csv python profiles/trace_one.py $c > gpurun_out/r2i/log$cfg.txt 2>&1
done
