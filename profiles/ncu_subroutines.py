#!/usr/bin/env python
"""Out-of-line subroutines of one profiled launch (everything that ends in RET.REL.NODEC): calls per warp, instructions per call,
and a guess at what it is (fp64 division / sqrt / rcp slow paths are compiler-generated and carry a wrong source line).
usage: ncu_subroutines.py prof.ncu-rep kernel_regex launch"""
import csv
import subprocess
import sys

rep, launch = sys.argv[1], int(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seg, segs, total = [], [], 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ie = int(r[ix['Instructions Executed']])
    except ValueError:
        continue
    total += ie
    s = r[ix['Source']]
    seg.append((s, ie))
    if 'RET.REL' in s or s.strip().startswith('EXIT') or 'BRA' in s and False:
        segs.append(seg)
        seg = []
print('total warp instructions %d' % total)
# a subroutine body = the instructions after the previous RET/EXIT up to this RET (approximation: contiguous layout)
for sg in segs:
    if 'RET.REL' not in sg[-1][0]:
        continue
    calls = sg[-1][1]
    if calls == 0:
        continue
    # walk back while execution counts stay in the same ballpark as the call count (body of the helper)
    body = []
    for s, ie in reversed(sg):
        if ie > 64 * max(calls, 1):
            break
        body.append((s, ie))
    inst = sum(ie for _, ie in body)
    txt = ' '.join(s for s, _ in body)
    kind = 'fp64 div slow path' if '8.98846567431157953865e+307' in txt and 'RCP64H' in txt else 'fp64 sqrt slow path' if 'RSQ64H' in txt else 'fp64 rcp slow path' if 'RCP64H' in txt else '?'
    nl = sum(ie for s, ie in body if 'STL' in s or 'LDL' in s)
    print('calls %10d  warp-instr %11d (%4.1f%% of launch)  local ld/st %10d  %s' % (calls, inst, 100. * inst / total, nl, kind))
