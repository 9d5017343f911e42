# entry step into the world taken inside the first k_nav
mkdir -p gpurun_out/r2x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2x/pytest.log 2>&1
tail -5 gpurun_out/r2x/pytest.log
python profiles/diff_modes.py 2 0 4000000; python profiles/diff_modes.py 5 20 2000000
for e in "RB_INIT=0" "RB_INIT=1" "RB_INIT=2"; do
for c in "1 0 9000000 3" "2 1 11115556 3" "3 0 9000000 3" "4 0 10000000 3" "5 20 10000000 3 rings=10"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-170 >> gpurun_out/r2x/survey.log
done
done
cat gpurun_out/r2x/survey.log
for e in "RB_INIT=0" "RB_INIT=1"; do
env $e timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 73 -c 12 --csv --log-file gpurun_out/r2x/metrics_cfg2_$e.csv python profiles/trace_one.py 2 1 11115556 2 > /dev/null 2>&1
python - "gpurun_out/r2x/metrics_cfg2_$e.csv" <<'PY'
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; iK = hdr.index('Kernel Name'); iM = hdr.index('Metric Name'); iV = hdr.index('Metric Value'); iI = hdr.index('ID')
cur = {}
for r in rows[1:]:
    cur.setdefault((r[iI], r[iK][:24]), {})[r[iM].split('__')[0][:4] + r[iM][-10:]] = r[iV]
for k, v in cur.items(): print(k, v)
PY
done
