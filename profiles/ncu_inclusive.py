#!/usr/bin/env python
"""Inclusive attribution of one profiled launch: executed warp instructions under every inlined frame (nvdisasm -gi gives the
whole inline chain of each SASS instruction), as a tree below the kernel body.
usage: ncu_inclusive.py prof.ncu-rep build/rb_trace_v_X.o kernel launch [min_percent]"""
import bisect
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.environ.get('RB_PROFILE_SRC') or os.path.join(HERE, '..', 'robast_b200', 'csrc')  # RB_PROFILE_SRC: the sources the profiled build came from


def func_table(path):
    starts = []
    for i, l in enumerate(open(path).read().splitlines()):
        m = re.search(r'(?:RB_HD|__device__|__global__)[^(]*?\b(\w+)\s*\(', l)
        if m and not l.lstrip().startswith('//'):
            starts.append((i + 1, m.group(1)))
    return starts


TABLES = {f: func_table(os.path.join(CSRC, f)) for f in ('rb_device.cuh', 'rb_trace_kernel.cuh')}


def fname(f, line):
    if f not in TABLES:
        return f
    t = TABLES[f]
    k = bisect.bisect_right([x[0] for x in t], line) - 1
    return t[k][1] if k >= 0 else f


def chains(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-gi', '-c', cubin], capture_output=True, text=True).stdout
    m, cur, insec, pending = {}, [], False, []
    for ln in out.splitlines():
        ms = re.match(r'^\s*\.section\s+(\S+)', ln)
        if ms:
            insec = ms.group(1).startswith('.text.') and kernel in ms.group(1).split(',')[0]
            continue
        if not insec:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if mm:
            if not pending or pending[-1][1] is None:
                pending = []
            pending.append(((os.path.basename(mm.group(1)), int(mm.group(2))), (os.path.basename(mm.group(3)), int(mm.group(4))) if mm.group(3) else None))
            if mm.group(3) is None:
                cur = [p[0] for p in pending]
            else:
                cur = [p[0] for p in pending] + [pending[-1][1]]
            continue
        mi = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);', ln)
        if mi:
            m[int(mi.group(1), 16)] = list(cur)
            pending = []
    return m


def main():
    rep, obj, kernel, launch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    minp = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    cm = chains(obj, kernel)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    tree = collections.Counter()
    base, tot = None, 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            addr, ie = int(r[ix['Address']], 16), int(r[ix['Instructions Executed']])
        except ValueError:
            continue
        if base is None:
            base = addr
        ch = cm.get(addr - base, [])
        names = []
        for f, l in reversed(ch):  # outermost first
            n = fname(f, l)
            if not names or names[-1] != n:
                names.append(n)
        tot += ie
        for k in range(1, len(names) + 1):
            tree[tuple(names[:k])] += ie
    print('total warp instructions', tot)
    for path in sorted(tree, key=lambda p: [(-tree[p[:k + 1]], p[k]) for k in range(len(p))]):
        v = tree[path]
        if 100. * v / tot >= minp:
            print('%s%-30s %5.1f%%' % ('  ' * (len(path) - 1), path[-1], 100. * v / tot))


if __name__ == '__main__':
    main()
