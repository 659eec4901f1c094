#!/usr/bin/env python3
"""Per source line of ONE kernel: executed warp instructions by opcode, from
`ncu -i rep --page source --csv --print-source sass,cuda` (kernels built with
-lineinfo).  Shows where the non-FP64 instructions of a loop come from.

    ncu -i prof.ncu-rep --page source --csv --print-source sass,cuda > src.csv
    python profiles/line_mix.py src.csv [min_share_percent]
"""
import csv
import sys
from collections import defaultdict


def main(path, min_share=0.4):
    rows = list(csv.reader(open(path)))
    cur_file, cur_line, cur_src = '?', 0, ''
    per_line = defaultdict(lambda: defaultdict(float))
    text = {}
    total = 0.0
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = {name: i for i, name in enumerate(r)}
            ex_col = r.index('Instructions Executed')
            continue
        if hdr is None or r[0] in ('Function Name',):
            continue
        if r[0] != '':                       # a CUDA source line
            cur_line, cur_src = int(r[0]), r[1].strip()
            text[(cur_file, cur_line)] = cur_src
            continue
        sass = r[3].strip()
        parts = sass.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
        op = op.rstrip(';').split('.')[0]
        try:
            n = float(r[ex_col] or 0)
        except ValueError:
            n = 0.0
        per_line[(cur_file, cur_line)][op] += n
        total += n
    print('total warp instructions %.0f' % total)
    for key in sorted(per_line, key=lambda k: -sum(per_line[k].values())):
        ops = per_line[key]
        tot = sum(ops.values())
        if 100 * tot / total < min_share:
            break
        mix = ' '.join('%s %.2f%%' % (o, 100 * c / total)
                       for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print('%5.2f%%  %s:%d  %s\n        %s' % (100 * tot / total, key[0], key[1],
                                                   text.get(key, '')[:100], mix))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4)
