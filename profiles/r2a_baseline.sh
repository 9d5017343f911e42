set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
nproc >> gpurun_out/r2a/smi.txt
timeout 600 python -m pytest tests/test_parity_branches.py -m gpu -x -q > gpurun_out/r2a/pytest_branches.log 2>&1
tail -3 gpurun_out/r2a/pytest_branches.log
for args in "1 0 9000000" "2 1 11115556" "3 0 9000000" "3 3 9000000" "4 0 10000000" "5 20 10000000 3 rings=10" "5 20 10000000 3 rings=10 precalc=1"; do
  timeout 300 python profiles/trace_one.py $args >> gpurun_out/r2a/survey.log 2>&1
done
cat gpurun_out/r2a/survey.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -c 2 -o gpurun_out/r2a/cfg4_kstep python profiles/trace_one.py 4 0 4000000 1 > gpurun_out/r2a/ncu4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -c 2 -o gpurun_out/r2a/cfg5_kstep python profiles/trace_one.py 5 20 4000000 1 rings=10 > gpurun_out/r2a/ncu5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -c 2 -o gpurun_out/r2a/cfg2_kstep python profiles/trace_one.py 2 1 11115556 1 > gpurun_out/r2a/ncu2.log 2>&1
ls -la gpurun_out/r2a
