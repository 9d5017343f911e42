mkdir -p gpurun_out/r3b
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nav|k_shade" -s 48 -c 2 -o gpurun_out/r3b/cfg5 python profiles/trace_one.py 5 20 10000000 1 rings=10 > gpurun_out/r3b/ncu5.log 2>&1
tail -2 gpurun_out/r3b/ncu5.log
ls -la gpurun_out/r3b
