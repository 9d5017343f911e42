mkdir -p gpurun_out/r4d
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 2 -o gpurun_out/r4d/cfg5 python profiles/trace_one.py 5 20 10000000 1 rings=10 > gpurun_out/r4d/ncu5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 52 -c 2 -o gpurun_out/r4d/cfg4 python profiles/trace_one.py 4 0 10000000 1 > gpurun_out/r4d/ncu4.log 2>&1
ls -la gpurun_out/r4d
