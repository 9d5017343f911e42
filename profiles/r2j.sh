mkdir -p gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py -m gpu -x -q > gpurun_out/r2j/pytest.log 2>&1
tail -3 gpurun_out/r2j/pytest.log
for c in "1 0 9000000 3" "2 1 11115556 3" "3 0 9000000 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  timeout 300 python profiles/trace_one.py $c 2>&1 | cut -c1-170 >> gpurun_out/r2j/survey.log
done
cat gpurun_out/r2j/survey.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_init" -c 1 -o gpurun_out/r2j/cfg5_init python profiles/trace_one.py 5 20 4000000 1 rings=10 > gpurun_out/r2j/ncu5.log 2>&1
