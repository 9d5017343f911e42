mkdir -p gpurun_out/r2h
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -c 2 -o gpurun_out/r2h/cfg2 python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2h/ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -c 2 -o gpurun_out/r2h/cfg5 python profiles/trace_one.py 5 20 4000000 1 rings=10 > gpurun_out/r2h/ncu5.log 2>&1
ls gpurun_out/r2h
