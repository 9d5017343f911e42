mkdir -p gpurun_out/r3n
timeout 1500 python bench.py > gpurun_out/r3n/bench.json 2> gpurun_out/r3n/bench.err
tail -c 300 gpurun_out/r3n/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3n/bench_ref.json 2> gpurun_out/r3n/bench_ref.err
tail -c 400 gpurun_out/r3n/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3n/launches_bench_davies_cotton.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-e2e > gpurun_out/r3n/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_nav|k_shade|k_trace|k_compact" -s 73 -c 6 --csv --log-file gpurun_out/r3n/metrics_cfg2.csv python profiles/trace_one.py 2 1 11115556 2 > /dev/null 2>&1
