#!/usr/bin/env python
"""Digest an .ncu-rep into the handful of numbers DESIGN.md/profiles quote.  Usage: ncu_digest.py file.ncu-rep [--sass]"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '?')[:90])
        for w in WANT:
            if w in d:
                print('  %-62s %s %s' % (w, d[w], units[hdr.index(w)]))


def sass(rep, launch=0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(launch), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot, ops, opi, n, ninst = collections.Counter(), collections.Counter(), collections.Counter(), 0, 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            s, ie = int(r[ix['# Samples']]), int(r[ix['Instructions Executed']])
        except ValueError:
            continue
        n += s
        ninst += 1
        for c in stall:
            tot[c] += int(r[ix[c]] or 0)
        m = r[ix['Source']].split()
        o = m[1] if m[0].startswith('@') else m[0]
        o = '.'.join(o.split('.')[:2]) if o.startswith(('MUFU', 'F2I', 'I2F', 'F2F')) else o.split('.')[0]
        ops[o] += s
        opi[o] += ie
    print('SASS instructions in kernel image: %d ; samples %d' % (ninst, n))
    print('stalls: ' + ', '.join('%s %.1f%%' % (k[6:], 100. * v / n) for k, v in tot.most_common(8)))
    ti = sum(opi.values())
    print('opcodes by executed instructions: ' + ', '.join('%s %.1f%%' % (k, 100. * v / ti) for k, v in opi.most_common(16)))
    print('opcodes by samples: ' + ', '.join('%s %.1f%%' % (k, 100. * v / n) for k, v in ops.most_common(12)))


if __name__ == '__main__':
    raw(sys.argv[1])
    if '--sass' in sys.argv:
        for l in range(int(sys.argv[sys.argv.index('--sass') + 1]) if len(sys.argv) > sys.argv.index('--sass') + 1 else 1):
            sass(sys.argv[1], l)
