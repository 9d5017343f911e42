mkdir -p gpurun_out/r3m
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3m/pytest.log 2>&1
tail -3 gpurun_out/r3m/pytest.log
python profiles/diff_modes.py 1 0 4000000; python profiles/diff_modes.py 3 0 4000000
for c in "1 0 9000000 3" "1 0 1000000 5" "3 0 25000000 3" "2 1 11115556 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  timeout 300 python profiles/trace_one.py $c 2>&1 | cut -c1-150 >> gpurun_out/r3m/survey.log
done
cat gpurun_out/r3m/survey.log
