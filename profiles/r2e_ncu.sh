mkdir -p gpurun_out/r2e
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bounce -c 1 -o gpurun_out/r2e/cfg2 python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2e/ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bounce -c 1 -o gpurun_out/r2e/cfg5 python profiles/trace_one.py 5 20 4000000 1 rings=10 > gpurun_out/r2e/ncu5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bounce -c 1 -o gpurun_out/r2e/cfg4 python profiles/trace_one.py 4 0 4000000 1 > gpurun_out/r2e/ncu4.log 2>&1
ls gpurun_out/r2e
