mkdir -p gpurun_out/r2c
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -o gpurun_out/r2c/cfg2 python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2c/ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -o gpurun_out/r2c/cfg5 python profiles/trace_one.py 5 20 4000000 1 rings=10 > gpurun_out/r2c/ncu5.log 2>&1
tail -2 gpurun_out/r2c/ncu2.log gpurun_out/r2c/ncu5.log
