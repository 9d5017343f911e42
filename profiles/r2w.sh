# local-box pre-test in nb_eval + k_locate (start nodes only) instead of k_init (copy of the rays)
mkdir -p gpurun_out/r2w
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2w/pytest.log 2>&1
tail -3 gpurun_out/r2w/pytest.log
python profiles/diff_modes.py 2 0 4000000; python profiles/diff_modes.py 5 20 2000000
for e in "RB_INIT=0" "RB_INIT=2" "RB_INIT=1"; do
for c in "1 0 9000000 3" "2 1 11115556 3" "3 0 9000000 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-170 >> gpurun_out/r2w/survey.log
done
done
cat gpurun_out/r2w/survey.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 73 -c 12 --csv --log-file gpurun_out/r2w/metrics_cfg2.csv python profiles/trace_one.py 2 1 11115556 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('gpurun_out/r2w/metrics_cfg2.csv') if l.startswith('"')))
hdr = rows[0]; iK = hdr.index('Kernel Name'); iM = hdr.index('Metric Name'); iV = hdr.index('Metric Value'); iI = hdr.index('ID')
cur = {}
for r in rows[1:]:
    cur.setdefault((r[iI], r[iK][:40]), {})[r[iM]] = r[iV]
for k, v in cur.items(): print(k, v)
PY
