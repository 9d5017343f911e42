mkdir -p gpurun_out/r3p
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 3 -o gpurun_out/r3p/cfg2 python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r3p/ncu2.log 2>&1
tail -2 gpurun_out/r3p/ncu2.log
