mkdir -p gpurun_out/r3e
RB_INIT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav" -s 26 -c 1 -o gpurun_out/r3e/cfg4_nav3 python profiles/trace_one.py 4 0 10000000 1 > gpurun_out/r3e/ncu4.log 2>&1
tail -2 gpurun_out/r3e/ncu4.log
