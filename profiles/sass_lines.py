#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines (needs the .o/.cubin the profiled build came from).

usage: sass_lines.py prof.ncu-rep build/rb_trace_v_X.o [launch_index] [top_n] [kernel name, default k_nav]
Joins `ncu --page source --print-source sass` (instructions executed, stall samples, by address) with
`nvdisasm -g` (address -> file:line) and prints the hottest source lines and functions."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


FUNCS = {}


def line_map(obj, kernel='k_nav'):
    """address -> (file, line) of the .text section of `kernel` only (each kernel's section starts at address 0)"""
    tmp = tempfile.mkdtemp()
    if obj.endswith('.cubin'):
        cubins = [obj]
    else:
        subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')]
    out = subprocess.run(['nvdisasm', '-g', '-c', cubins[0]], capture_output=True, text=True).stdout
    m, cur, fn = {}, ('?', 0), '?'
    FUNCS.clear()
    insec = False
    for ln in out.splitlines():
        ms = re.match(r'^\s*\.section\s+(\S+)', ln)
        if ms:
            insec = ms.group(1).startswith('.text.') and kernel in ms.group(1).split(',')[0]
            continue
        if not insec:
            continue
        mf = re.match(r'^(\S+):\s*$', ln)
        if mf and not mf.group(1).startswith('.L'):
            fn = re.sub(r'^\$?_Z\dk_(trace|step)[^$]*\$', '', mf.group(1))
            fn = re.sub(r'^\.text\..*', kernel + '(body)', fn)
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);', ln)
        if mm:
            m[int(mm.group(1), 16)] = cur
            FUNCS[int(mm.group(1), 16)] = fn
    return m


def main():
    rep, obj = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(obj, sys.argv[5] if len(sys.argv) > 5 else 'k_nav')
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    base = None
    per_line = collections.defaultdict(lambda: [0, 0, 0])  # inst, samples, no_inst samples
    per_fn = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            addr, ie, s = int(r[ix['Address']], 16), int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
        except ValueError:
            continue
        if base is None:
            base = addr
        key = lm.get(addr - base, ('?', 0))
        a = per_line[key]
        a[0] += ie
        a[1] += s
        a[2] += int(r[ix['stall_no_inst']] or 0)
        b = per_fn[FUNCS.get(addr - base, '?')]
        b[0] += ie
        b[1] += s
        b[2] += int(r[ix['stall_no_inst']] or 0)
        tot_i += ie
        tot_s += s
    print('kernel: %s   executed warp-instructions %d, samples %d' % (rows[0][1][:80], tot_i, tot_s))
    print('%-70s %7s %7s %7s' % ('function', 'inst%', 'smpl%', 'noinst%'))
    for f, (ie, sm, ni) in sorted(per_fn.items(), key=lambda kv: -kv[1][1]):
        print('%-70s %6.2f%% %6.2f%% %6.2f%%' % (f[:70], 100. * ie / tot_i, 100. * sm / tot_s, 100. * ni / max(1, sm)))
    src_cache = {}

    def src(f, l):
        if f not in src_cache:
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'robast_b200', 'csrc', f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[f]
        return L[l - 1].strip()[:90] if 0 < l <= len(L) else ''
    print('%-28s %7s %7s %7s  %s' % ('file:line', 'inst%', 'smpl%', 'noinst%', 'source'))
    for (f, l), (ie, s, ni) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-28s %6.2f%% %6.2f%% %6.2f%%  %s' % ('%s:%d' % (f, l), 100. * ie / tot_i, 100. * s / tot_s, 100. * ni / max(1, s), src(f, l)))


if __name__ == '__main__':
    main()
