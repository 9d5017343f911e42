mkdir -p gpurun_out/r2y
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 4 -o gpurun_out/r2y/cfg2 python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2y/ncu2.log 2>&1
tail -3 gpurun_out/r2y/ncu2.log
ls -la gpurun_out/r2y
