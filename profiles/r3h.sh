# step limit into intersection operands; L1 carve-out preference; bench clock sampling after queueing
mkdir -p gpurun_out/r3h
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r3h/pytest.log 2>&1
tail -3 gpurun_out/r3h/pytest.log
for e in "RB_L1_CARVEOUT=0" "RB_L1_CARVEOUT=-1" "RB_L1_CARVEOUT=25"; do
for c in "1 0 9000000 3" "2 1 11115556 3" "3 0 9000000 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-150 >> gpurun_out/r3h/survey.log
done
done
cat gpurun_out/r3h/survey.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-configs --no-cpu --no-e2e > gpurun_out/r3h/b.json 2> gpurun_out/r3h/b.err
python -c "
import json
d = json.loads(open('gpurun_out/r3h/b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['step_ms'], d['clocks'])"
