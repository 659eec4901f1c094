#!/usr/bin/env python3
"""Merges gpurun_out/configs_multi_<N>gpu.json (tools/bench_configs_multi.py at
N = 1, 2, 4, 8) into one markdown table: one row per configuration and N.

    python tools/configs_table.py > profiles/r02_configs_8gpu.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    runs = {}
    for n in (1, 2, 4, 8):
        path = os.path.join(ROOT, 'gpurun_out', 'configs_multi_%dgpu.json' % n)
        if os.path.exists(path):
            runs[n] = json.load(open(path))
    if not runs:
        sys.exit('no gpurun_out/configs_multi_*gpu.json found')
    ns = sorted(runs)
    print('# BASELINE.json configurations on 1 / 2 / 4 / 8 B200 (tools/bench_configs_multi.py)\n')
    print('Device-timed, median of the repetitions, maximum over ranks; measured FP64 peak per GPU: '
          + ', '.join('%.1f TFLOP/s (N=%d)' % (runs[n]['fp64_peak_tflops_measured'], n) for n in ns) + '.\n')
    print('## configs[4]: FitzHugh-Nagumo population sweep, GLOBAL population split over N GPUs '
          '(strong scaling), every rank ends up with all scores\n')
    print('| global population | ' + ' | '.join('N=%d: ms, evals/s, %% of FP64 peak per GPU' % n for n in ns) + ' |')
    print('|---|' + '---|' * len(ns))
    sizes = [r['n_global'] for r in runs[ns[0]]['fhn_sweep']]
    for k, size in enumerate(sizes):
        cells = []
        for n in ns:
            sw = runs[n]['fhn_sweep']
            if k < len(sw):
                r = sw[k]
                cells.append('%.4f, %.3g, %.1f %%' % (r['ms'], r['evals_per_s'],
                                                      100 * r['frac_of_peak_per_gpu']))
            else:
                cells.append('-')
        print('| %d | ' % size + ' | '.join(cells) + ' |')
    print('\nSpeed-up over one GPU at the same global population:\n')
    print('| global population | ' + ' | '.join('N=%d' % n for n in ns) + ' |')
    print('|---|' + '---|' * len(ns))
    for k, size in enumerate(sizes):
        base = runs[ns[0]]['fhn_sweep'][k]['ms']
        print('| %d | ' % size + ' | '.join(
            '%.2f' % (base / runs[n]['fhn_sweep'][k]['ms']) if k < len(runs[n]['fhn_sweep']) else '-'
            for n in ns) + ' |')
    print('\n## The other configurations\n')
    print('| configuration | ' + ' | '.join('N=%d' % n for n in ns) + ' |')
    print('|---|' + '---|' * len(ns))
    rows = [
        ('configs[1]: xNES generation (ask + score + exchange + tell), population 65 536 in total, '
         'DeviceXNES: ms per generation', lambda r: '%.3f' % r['xnes']['ms_per_generation']),
        ('configs[2]: HaarioBardenetACMC, 16 384 chains on HH-IK (4800 points), chains sharded, no '
         'exchange: ms per iteration (projected s for 10^4 iterations)',
         lambda r: '%.4f (%.2f s)' % (r['hh_ik_hbac']['ms_per_iteration'],
                                      r['hh_ik_hbac']['projected_s_for_1e4_iterations'])),
        ('configs[3]: DifferentialEvolutionMCMC, 4096 chains on Lotka-Volterra, chains sharded, '
         'positions all-gathered per iteration: ms per iteration',
         lambda r: '%.4f' % r['lv_de']['ms_per_iteration']),
        ('configs[3]: the same on SIR: ms per iteration',
         lambda r: '%.4f' % r['sir_de']['ms_per_iteration']),
    ]
    for label, fmt in rows:
        print('| %s | ' % label + ' | '.join(fmt(runs[n]) for n in ns) + ' |')


if __name__ == '__main__':
    main()
