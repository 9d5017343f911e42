mkdir -p gpurun_out/r4c
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_scene_cache.py -m gpu -x -q > gpurun_out/r4c/pytest.log 2>&1
tail -3 gpurun_out/r4c/pytest.log
python profiles/diff_modes.py 1 0 4000000
for c in "1 0 9000000 3" "1 0 1000000 5" "1 0.3 1000000 5" "3 0 25000000 3"; do
  timeout 300 python profiles/trace_one.py $c 2>&1 | cut -c1-150 >> gpurun_out/r4c/survey.log
done
cat gpurun_out/r4c/survey.log
