#!/usr/bin/env python
"""Headline metrics of every kernel in an ncu report (development aid).  usage: ncu_brief.py x.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
for r in rows[2:]:
    print("==", r[h.index('Kernel Name')][:90])
    for k in want:
        if k in h:
            print(f"  {k:62s} {r[h.index(k)]:>16s} {u[h.index(k)]}")
    st = sorted(((float(r[i] or 0), x.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                 for i, x in enumerate(h) if 'issue_stalled' in x and 'per_issue_active' in x), reverse=True)
    print("  stalls: " + ", ".join(f"{n} {v:.1f}" for v, n in st if v > 0.3))
