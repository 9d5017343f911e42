mkdir -p gpurun_out/r3y
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r3y/pytest.log 2>&1
tail -3 gpurun_out/r3y/pytest.log
python profiles/diff_modes.py 5 20 4000000 rings=10
for e in "RB_COOP=1" "RB_COOP=0"; do
for c in "5 20 10000000 3 rings=10" "5 0 10000000 3 rings=10" "5 20 10000000 3 rings=30"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-170 >> gpurun_out/r3y/survey.log
done
done
cat gpurun_out/r3y/survey.log
