#!/usr/bin/env python
"""Which source lines call the compiler's fp64 division / sqrt slow-path subroutines, and how often (per launch).
usage: ncu_slowpath_callers.py prof.ncu-rep build/rb_trace_v_X.o kernel launch"""
import collections
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_lines as SL

rep, obj, kernel, launch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
lm = SL.line_map(obj, kernel)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ins.append((int(r[ix['Address']], 16), r[ix['Source']], int(r[ix['Instructions Executed']])))
    except ValueError:
        pass
base = ins[0][0]
warps = max(ie for _, _, ie in ins[:200])
# slow-path helpers: short RET-terminated runs containing MUFU.RCP64H / RSQ64H whose every instruction runs ~ once per call
helpers = {}
start = 0
for i, (a, s, ie) in enumerate(ins):
    if 'RET.REL' in s:
        j = i
        while j > 0 and i - j < 200 and 'RET.REL' not in ins[j - 1][1] and 'EXIT' not in ins[j - 1][1] and not ins[j - 1][1].strip().startswith('BRA') or (j > 0 and i - j < 200 and ins[j - 1][1].strip().startswith('BRA') and ins[j - 1][2] <= ie):
            j -= 1
        body = ' '.join(x[1] for x in ins[j:i + 1])
        if i - j < 200 and ('RCP64H' in body or 'RSQ64H' in body):
            kind = 'div' if '8.98846567431157953865e+307' in body else 'sqrt' if 'RSQ64H' in body else 'rcp'
            for k in range(j, i + 1):
                helpers[ins[k][0]] = (kind, ins[j][0])
src = open(os.path.join(os.environ.get('RB_PROFILE_SRC') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'robast_b200', 'csrc'), 'rb_device.cuh')).read().splitlines()
by = collections.Counter()
for a, s, ie in ins:
    m = re.search(r'CALL\.REL\.NOINC\s+0x([0-9a-f]+)', s)
    if m and ie > 0:
        t = int(m.group(1), 16)
        if t in helpers:
            f, l = lm.get(a - base, ('?', 0))
            by[(helpers[t][0], f, l)] += ie
print('warps in the launch ~ %d' % warps)
for (kind, f, l), v in by.most_common(40):
    text = src[l - 1].strip()[:100] if f == 'rb_device.cuh' and 0 < l <= len(src) else ''
    print('%-4s calls %9d (%5.2f / warp)  %s:%d  %s' % (kind, v, v / warps, f[3:], l, text))
