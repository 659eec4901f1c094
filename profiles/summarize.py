#!/usr/bin/env python3
"""Turn an `ncu --set full` report into the text summary committed under
profiles/.  Usage:

    python profiles/summarize.py gpurun_out/prof_<tag>.ncu-rep [kernel-regex] > profiles/<tag>_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__occupancy_limit_registers', 'occupancy limit: registers (blocks)'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit: shared memory (blocks)'),
    ('sm__warps_active.avg.per_cycle_active', 'warps active per SM'),
    ('smsp__warps_active.avg.per_cycle_active', 'warps active per scheduler'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'warps eligible per scheduler'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU/F2F) pipe %'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads per instruction (of 32)'),
    ('smsp__average_warps_active_per_inst_executed.ratio', 'cycles between issues per warp'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall: wait (fixed latency)'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall: short scoreboard (MUFU/LDS)'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall: math pipe throttle'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall: not selected'),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall: branch resolving'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall: long scoreboard (global)'),
    ('dram__bytes_read.sum', 'DRAM bytes read'),
    ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts'),
]


def main():
    rep = sys.argv[1]
    regex = sys.argv[2] if len(sys.argv) > 2 else None
    cmd = ['ncu', '-i', rep, '--page', 'raw', '--csv']
    if regex:
        cmd += ['--kernel-name', 'regex:' + regex]
    text = subprocess.run(cmd, stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    print('# ncu summary of `%s`\n' % rep.split('/')[-1])
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print('## %s\n' % name[:140])
        print('| metric | value | unit |')
        print('|---|---|---|')
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print('| %s (`%s`) | %s | %s |' % (label, key, r[i], units[i]))
        print()


if __name__ == '__main__':
    main()
