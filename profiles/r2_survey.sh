# usage: bash profiles/r2_survey.sh TAG [pytest-args]   -> gpurun_out/TAG/{pytest.log,survey.log}
TAG=$1; shift
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py -m gpu -x -q "$@" > gpurun_out/$TAG/pytest.log 2>&1
tail -5 gpurun_out/$TAG/pytest.log
for args in "1 0 9000000" "2 1 11115556" "3 0 9000000" "4 0 10000000" "5 20 10000000 3 rings=10" "5 20 10000000 3 rings=10 precalc=1"; do
  timeout 300 python profiles/trace_one.py $args >> gpurun_out/$TAG/survey.log 2>&1
done
cat gpurun_out/$TAG/survey.log
