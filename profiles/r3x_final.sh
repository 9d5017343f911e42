mkdir -p gpurun_out/r3x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r3x/pytest_gpu.log 2>&1
tail -3 gpurun_out/r3x/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r3x/bench.json 2> gpurun_out/r3x/bench.err
tail -c 200 gpurun_out/r3x/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3x/bench_ref.json 2> gpurun_out/r3x/bench_ref.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3x/bench.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.2f step_ms %s e2e %.4g launches %d' % (d['value'], d['ms_per_step'], d['step_ms'], d['e2e']['value'], d['gpu_launches']))
for k, c in d['configs'].items():
    print(k, '%.4g rays/s %.3f ms, e2e %.3g, 1k %.0f us' % (c['value'], c['ms_per_trace'], c['e2e']['value'], c['latency_us_1k_rays']), c.get('value_with_precalculated_tmm_table'))
print(d['cfg5_strong']['value'], d['cfg5_strong']['ms_per_step'], d['cpu_baseline']['value'])
r = json.loads(open('gpurun_out/r3x/bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
