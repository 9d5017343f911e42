# usage: bash profiles/r2_quick.sh TAG  -> gpurun_out/TAG/survey.log  (throughput of the five configs, lock-step and free-running)
TAG=$1; shift
mkdir -p gpurun_out/$TAG
for args in "1 0 9000000" "2 1 11115556" "3 0 9000000" "4 0 10000000" "5 20 10000000 3 rings=10"; do
  timeout 300 python profiles/trace_one.py $args >> gpurun_out/$TAG/survey.log 2>&1
  RB_NO_LOCKSTEP=1 timeout 300 python profiles/trace_one.py $args 2>&1 | sed 's/^/NOLOCK /' >> gpurun_out/$TAG/survey.log
done
cut -c1-200 gpurun_out/$TAG/survey.log
