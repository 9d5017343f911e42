mkdir -p gpurun_out/r2s
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s/launches_bench_davies_cotton.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-e2e > gpurun_out/r2s/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_init|k_nav|k_shade|k_compact" -s 73 -c 4 -o gpurun_out/r2s/cfg2 python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2s/ncu2.log 2>&1
ls -la gpurun_out/r2s
