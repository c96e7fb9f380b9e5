#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics that decide what bounds a kernel.  Usage: ncu_keys.py <raw.csv>"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
def col(name):
    i = [j for j, h in enumerate(hdr) if h == name][0]
    return [r[i] for r in rows[2:]]
print([k[5:45] for k in col('Kernel Name')])
for n in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
          'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
          'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']:
    try:
        print(n[-62:].rjust(62), col(n))
    except Exception:
        print('missing', n)
