mkdir -p gpurun_out/r3z
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "not tutorial" > gpurun_out/r3z/pytest.log 2>&1
tail -2 gpurun_out/r3z/pytest.log
for c in "1 0 300000 1" "2 1 300000 1" "3 0 300000 1" "5 20 300000 1 rings=10"; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/trace_one.py $c > gpurun_out/r3z/san.log 2>&1
  echo "cfg $c -> exit $?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r3z/san.log; grep "ERROR SUMMARY" gpurun_out/r3z/san.log | tail -1
done
