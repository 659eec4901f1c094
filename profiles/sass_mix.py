#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump of ONE kernel by SASS opcode:
executed warp-instructions, average active threads, stall samples.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv
    python profiles/sass_mix.py src.csv [top]
"""
import csv
import sys
from collections import defaultdict


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    # the dump may hold several kernels; use the first block only
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    col = {name: i for i, name in enumerate(hdr)}
    ex = defaultdict(float)
    thr = defaultdict(float)
    smp = defaultdict(float)
    total = 0.0
    for r in rows[hdr_i + 1:]:
        if not r or r[0] == 'Kernel Name' or r[0] == 'Address':
            break
        src = r[col['Source']].strip()
        parts = src.split()
        op = parts[1] if parts and parts[0].startswith('@') else parts[0]
        op = op.rstrip(';')
        base = op.split('.')[0]
        n = float(r[col['Instructions Executed']] or 0)
        ex[base] += n
        thr[base] += float(r[col['Thread Instructions Executed']] or 0)
        smp[base] += float(r[col['# Samples']] or 0)
        total += n
    tot_s = sum(smp.values()) or 1.0
    print('%-10s %14s %7s %8s %8s' % ('opcode', 'warp-insts', 'share', 'threads',
                                      'samples'))
    for k in sorted(ex, key=lambda k: -ex[k])[:top]:
        print('%-10s %14.0f %6.1f%% %8.1f %7.1f%%' % (
            k, ex[k], 100 * ex[k] / total, thr[k] / max(ex[k], 1),
            100 * smp[k] / tot_s))
    print('%-10s %14.0f' % ('TOTAL', total))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
