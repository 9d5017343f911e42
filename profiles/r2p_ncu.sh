mkdir -p gpurun_out/r2p
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_nav" -s 24 -c 2 -o gpurun_out/r2p/cfg2_nav python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2p/ncu2.log 2>&1
ls -la gpurun_out/r2p
