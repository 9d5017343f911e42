mkdir -p gpurun_out/r3v
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r3v/pytest.log 2>&1
tail -3 gpurun_out/r3v/pytest.log
for c in "2 1 11115556 3" "4 0 10000000 3" "5 20 10000000 3 rings=10" "5 20 10000000 3 rings=10 precalc=1"; do
  timeout 300 python profiles/trace_one.py $c 2>&1 | cut -c1-150 >> gpurun_out/r3v/survey.log
done
cat gpurun_out/r3v/survey.log
