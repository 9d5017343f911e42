# light per-launch metric pass over every kernel of one timed trace of configs 2, 4, 5 (csv only)
mkdir -p gpurun_out/r2v
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,launch__registers_per_thread
for c in "2 1 11115556 1" "4 0 10000000 1" "5 20 10000000 1 rings=10"; do
  cfg=${c%% *}
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2v/metrics_cfg$cfg.csv python profiles/trace_one.py $c > gpurun_out/r2v/log$cfg.txt 2>&1
done
ls -la gpurun_out/r2v
