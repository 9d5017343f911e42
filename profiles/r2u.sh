mkdir -p gpurun_out/r2u
RB_SPLIT_EVAL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py -m gpu -x -q > gpurun_out/r2u/pytest.log 2>&1
tail -3 gpurun_out/r2u/pytest.log
RB_SPLIT_EVAL=1 python profiles/diff_modes.py 2 0 4000000; RB_SPLIT_EVAL=1 python profiles/diff_modes.py 5 20 2000000
for e in "RB_SPLIT_EVAL=0" "RB_SPLIT_EVAL=1"; do
for c in "1 0 9000000 3" "2 1 11115556 3" "3 0 9000000 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-170 >> gpurun_out/r2u/survey.log
done
done
cat gpurun_out/r2u/survey.log
