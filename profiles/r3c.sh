# coherent TMM as a column-vector recurrence
mkdir -p gpurun_out/r3c
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_branches.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r3c/pytest.log 2>&1
tail -3 gpurun_out/r3c/pytest.log
for e in "RB_INIT=1"; do
for c in "2 1 11115556 3" "4 0 10000000 3" "5 20 10000000 3 rings=10" "5 20 10000000 3 rings=10 precalc=1"; do
  env $e timeout 300 python profiles/trace_one.py $c 2>&1 | sed "s/^/$e /" | cut -c1-170 >> gpurun_out/r3c/survey.log
done
done
cat gpurun_out/r3c/survey.log
for c in "5 20 10000000 2 rings=10"; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_nav|k_shade|k_locate|k_trace" -s 49 -c 8 --csv --log-file gpurun_out/r3c/m.csv python profiles/trace_one.py $c > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('gpurun_out/r3c/m.csv') if l.startswith('"')))
hdr = rows[0]; iK = hdr.index('Kernel Name'); iM = hdr.index('Metric Name'); iV = hdr.index('Metric Value'); iI = hdr.index('ID')
cur = {}
for r in rows[1:]:
    cur.setdefault((int(r[iI]), r[iK][:14]), {})[r[iM]] = float(r[iV].replace(',',''))
for k, v in sorted(cur.items()):
    print(k, 'ms %.2f inst %.0fM rd %.2f GB wr %.2f GB' % (v['gpu__time_duration.sum']/1e6, v['smsp__inst_executed.sum']/1e6, v['dram__bytes_read.sum']/1e9, v['dram__bytes_write.sum']/1e9))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 24 -c 1 -o gpurun_out/r3c/cfg5 python profiles/trace_one.py 5 20 10000000 1 rings=10 > gpurun_out/r3c/ncu5.log 2>&1
