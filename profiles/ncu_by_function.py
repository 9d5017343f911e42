#!/usr/bin/env python
"""Executed warp instructions and stall samples of one profiled launch, summed per device function of rb_device.cuh /
rb_trace_kernel.cuh (innermost inlined frame, by line info).
usage: ncu_by_function.py prof.ncu-rep build/rb_trace_v_X.o [kernel name] [launch index]"""
import bisect
import collections
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_lines as SL

HERE = os.path.dirname(os.path.abspath(__file__))


def func_table(path):
    src = open(path).read().splitlines()
    starts = []
    for i, l in enumerate(src):
        m = re.search(r'(?:RB_HD|__device__|__global__)[^(]*?\b(\w+)\s*\(', l)
        if m and not l.lstrip().startswith('//'):
            starts.append((i + 1, m.group(1)))
    return starts


def main():
    rep, obj = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else 'k_nav'
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lm = SL.line_map(obj, kernel)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    tables = {f: func_table(os.path.join(os.environ.get('RB_PROFILE_SRC') or os.path.join(HERE, '..', 'robast_b200', 'csrc'), f)) for f in ('rb_device.cuh', 'rb_trace_kernel.cuh')}
    inst, smp = collections.Counter(), collections.Counter()
    base = None
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            addr = int(r[ix['Address']], 16) if 'Address' in ix else None
            ie, s = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
        except ValueError:
            continue
        f = '?'
        if base is None:
            base = addr
        addr -= base
        if addr in lm:
            fn, line = lm[addr]
            if fn in tables:
                t = tables[fn]
                k = bisect.bisect_right([x[0] for x in t], line) - 1
                f = t[k][1] if k >= 0 else fn
            else:
                f = fn
        inst[f] += ie
        smp[f] += s
    ti, ts = sum(inst.values()), sum(smp.values())
    print('executed warp instructions %d, samples %d' % (ti, ts))
    for f, v in inst.most_common(40):
        print('%-28s inst %5.1f%%  samples %5.1f%%' % (f, 100. * v / ti, 100. * smp[f] / ts))


if __name__ == '__main__':
    main()
