#!/usr/bin/env python
"""Static SASS census of a bounce-kernel object: instructions per device function of rb_device.cuh (by line info, innermost
inlined frame), spill traffic and the local-memory instruction count.  usage: sass_static.py build/rb_trace_v_X.o [top_n]"""
import bisect
import collections
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    obj, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    src = open(os.path.join(HERE, '..', 'robast_b200', 'csrc', 'rb_device.cuh')).read().splitlines()
    starts = []
    for i, l in enumerate(src):
        m = re.search(r'RB_HD[^(]*?\b(\w+)\s*\(', l)
        if m and not l.lstrip().startswith('//') and ('inline' in l or 'RB_NOINLINE' in l or 'static' in l or l.startswith('template')):
            starts.append((i + 1, m.group(1)))
    ln = [s[0] for s in starts]
    per_sec = collections.defaultdict(collections.Counter)
    ops = collections.defaultdict(collections.Counter)
    sec, cur = None, None
    for l in out.splitlines():
        ms = re.match(r'^\s*\.section\s+(\.text\.\S+?),', l)
        if ms:
            sec = ms.group(1)[6:]
            continue
        if l.startswith('\t.section') or re.match(r'^\s*\.section', l):
            sec = None if '.text.' not in l else sec
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mi = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(@!?\S+\s+)?(\S+)', l)
        if mi and sec:
            f = '?'
            if cur:
                if cur[0] == 'rb_device.cuh':
                    k = bisect.bisect_right(ln, cur[1]) - 1
                    f = starts[k][1] if k >= 0 else '?'
                else:
                    f = cur[0]
            per_sec[sec][f] += 1
            ops[sec][mi.group(2).split('.')[0]] += 1
    for s, c in per_sec.items():
        tot = sum(c.values())
        name = subprocess.run(['c++filt', s], capture_output=True, text=True).stdout.strip()[:110]
        print('== %s: %d SASS instructions; LDL %d STL %d; DFMA+DMUL+DADD %d' % (name, tot, ops[s]['LDL'], ops[s]['STL'], ops[s]['DFMA'] + ops[s]['DMUL'] + ops[s]['DADD']))
        if tot > 3000:
            print('   ' + ', '.join('%s %d' % kv for kv in c.most_common(top)))


if __name__ == '__main__':
    main()
