#!/usr/bin/env python
"""Executed warp instructions of one profiled launch per source line (innermost inlined frame), hottest first, and the
number of times the line's first instruction ran per warp (a call count).
usage: ncu_by_line.py prof.ncu-rep build/rb_trace_v_X.o kernel launch [top]"""
import collections
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_lines as SL

HERE = os.path.dirname(os.path.abspath(__file__))
rep, obj, kernel, launch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
lm = SL.line_map(obj, kernel)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst, first, thr = collections.Counter(), {}, collections.Counter()
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        addr = int(r[ix['Address']], 16)
        ie = int(r[ix['Instructions Executed']])
        te = int(r[ix['Thread Instructions Executed']])
    except (ValueError, KeyError):
        continue
    if base is None:
        base = addr
    key = lm.get(addr - base, ('?', 0))
    inst[key] += ie
    thr[key] += te
    first[key] = max(first.get(key, 0), ie)
tot = sum(inst.values())
src = {f: open(os.path.join(os.environ.get('RB_PROFILE_SRC') or os.path.join(HERE, '..', 'robast_b200', 'csrc'), f)).read().splitlines() for f in ('rb_device.cuh', 'rb_trace_kernel.cuh')}
print('total warp instructions', tot)
for (f, l), v in inst.most_common(top):
    text = src[f][l - 1].strip()[:110] if f in src and 0 < l <= len(src[f]) else ''
    print('%5.2f%% max/instr %9d lanes %4.1f %s:%d  %s' % (100. * v / tot, first[(f, l)], thr[(f, l)] / max(v, 1), f.replace('rb_', '')[:12], l, text))
