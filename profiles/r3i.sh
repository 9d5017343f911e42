mkdir -p gpurun_out/r3i
for i in 1 2 3 4; do
RB_BENCH_DIAG=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-configs --no-cpu --no-e2e > gpurun_out/r3i/b$i.json 2> gpurun_out/r3i/b$i.err
python -c "
import json
d = json.loads(open('gpurun_out/r3i/b$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['step_ms'])"
done
grep -h diag gpurun_out/r3i/b*.err | head -150 > gpurun_out/r3i/diag.txt
