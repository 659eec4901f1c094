#!/usr/bin/env python3
"""Per CUDA source line of one kernel: warp-stall samples and executed warp
instructions, from `ncu -i rep --page source --csv --print-source sass,cuda`
(kernels built with -lineinfo).  For latency-bound kernels: where a warp waits.

    python profiles/stall_lines.py src.csv [top]
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur = ('?', 0)
    text = {}
    samples = defaultdict(float)
    executed = defaultdict(float)
    by_op = defaultdict(lambda: defaultdict(float))
    hdr = None
    cur_file = '?'
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = {name: i for i, name in enumerate(r)}
            continue
        if hdr is None or r[0] == 'Function Name':
            continue
        if r[0] != '':
            cur = (cur_file, int(r[0]))
            text[cur] = r[1].strip()
            continue
        def col(name):
            i = hdr.get(name)
            try:
                return float(r[i]) if i is not None and r[i] != '' else 0.0
            except ValueError:
                return 0.0
        s = col('Warp Stall Sampling (All Samples)') or col('# Samples')
        e = col('Instructions Executed')
        samples[cur] += s
        executed[cur] += e
        op = r[3].strip().split()
        if op:
            name = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
            by_op[cur][name.split('.')[0]] += s
    total = sum(samples.values()) or 1.0
    print('total stall samples %.0f, executed warp instructions %.0f' % (total, sum(executed.values())))
    for key in sorted(samples, key=samples.get, reverse=True)[:top]:
        ops = sorted(by_op[key].items(), key=lambda kv: -kv[1])[:4]
        print('%5.1f %%  %9.0f instr  %s:%d  %s   [%s]' % (
            100 * samples[key] / total, executed[key], key[0], key[1], text.get(key, '')[:90],
            ', '.join('%s %.0f' % kv for kv in ops)))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
