mkdir -p gpurun_out/r2s
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 2 -o gpurun_out/r2s/cfg4 python profiles/trace_one.py 4 0 4000000 1 > gpurun_out/r2s/ncu4.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 2 -o gpurun_out/r2s/cfg5 python profiles/trace_one.py 5 20 4000000 1 rings=10 > gpurun_out/r2s/ncu5.log 2>&1
ls -la gpurun_out/r2s
