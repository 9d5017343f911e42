mkdir -p gpurun_out/r3g
for i in 1 2 3; do
timeout 600 python bench.py --steps 12 --warmup 3 --no-configs --no-cpu --no-e2e > gpurun_out/r3g/b$i.json 2> gpurun_out/r3g/b$i.err
python -c "
import json
d = json.loads(open('gpurun_out/r3g/b$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['step_ms'])"
done
