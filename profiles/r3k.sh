mkdir -p gpurun_out/r3k
for c in "1 0 9000000 2" "3 0 25000000 2"; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum --clock-control none -k regex:"k_nav|k_shade|k_trace|k_compact" -s 73 -c 9 --csv --log-file gpurun_out/r3k/m.csv python profiles/trace_one.py $c > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('gpurun_out/r3k/m.csv') if l.startswith('"')))
hdr = rows[0]; iK = hdr.index('Kernel Name'); iM = hdr.index('Metric Name'); iV = hdr.index('Metric Value'); iI = hdr.index('ID')
cur = {}
for r in rows[1:]:
    cur.setdefault((int(r[iI]), r[iK][:14]), {})[r[iM]] = float(r[iV].replace(',',''))
for k, v in sorted(cur.items()):
    ms = v['gpu__time_duration.sum']/1e6
    print(k, 'ms %.3f inst %.0fM rd %.2f GB wr %.2f GB  dram %.2f TB/s fp64 %.0f%% issue %.0f%% local ld/st %.0fM/%.0fM' % (ms, v['smsp__inst_executed.sum']/1e6, v['dram__bytes_read.sum']/1e9, v['dram__bytes_write.sum']/1e9, (v['dram__bytes_read.sum']+v['dram__bytes_write.sum'])/1e9/ms, v['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'], v['smsp__issue_active.avg.pct_of_peak_sustained_active'], v['smsp__sass_inst_executed_op_local_ld.sum']/1e6, v['smsp__sass_inst_executed_op_local_st.sum']/1e6))
PY
done
