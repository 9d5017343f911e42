mkdir -p gpurun_out/r3f
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r3f/pytest_gpu.log 2>&1
tail -3 gpurun_out/r3f/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r3f/bench.json 2> gpurun_out/r3f/bench.err
tail -c 600 gpurun_out/r3f/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3f/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'gpu_launches', 'clocks') if k in d})
for k, c in d.get('configs', {}).items():
    print(k, {q: c.get(q) for q in ('rays', 'value', 'ms_per_trace', 'latency_us_1k_rays')}, c.get('e2e', {}).get('value'), [ (r.get('kind'), r.get('value')) for r in c.get('cpu', [])] if isinstance(c.get('cpu'), list) else c.get('cpu'))
print(d.get('cfg5_strong'))
print(d.get('cpu_baseline'))
PY
