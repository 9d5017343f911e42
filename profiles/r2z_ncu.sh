mkdir -p gpurun_out/r2z
# cfg5: first call plans 24 rounds (24 k_nav + 24 k_shade); the timed call's first bounce is launches 48, 49
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 2 -o gpurun_out/r2z/cfg5 python profiles/trace_one.py 5 20 10000000 1 rings=10 > gpurun_out/r2z/ncu5.log 2>&1
tail -2 gpurun_out/r2z/ncu5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 2 -o gpurun_out/r2z/cfg4 python profiles/trace_one.py 4 0 10000000 1 > gpurun_out/r2z/ncu4.log 2>&1
tail -2 gpurun_out/r2z/ncu4.log
ls -la gpurun_out/r2z
