mkdir -p gpurun_out/r3j
for i in 1 2 3 4 5 6 7 8; do
RB_BENCH_DIAG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu --no-e2e > gpurun_out/r3j/b$i.json 2> gpurun_out/r3j/b$i.err
python -c "
import json
d = json.loads(open('gpurun_out/r3j/b$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['step_ms'])"
grep "containment call" gpurun_out/r3j/b$i.err | awk '{print \$5}' | tr '\n' ' '; echo
done
