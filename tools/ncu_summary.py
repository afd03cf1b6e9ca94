#!/usr/bin/env python
"""Markdown summary of an ncu report (--set full) for profiles/: per kernel the headline metrics,
the stall picture and the DRAM traffic; also rewrites profiles/traffic.json for the step kernel.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "<title>" "<command>" [--no-traffic] > profiles/rNN_ncu_kernels.md

--no-traffic: leave profiles/traffic.json alone (captures of other workloads than bench.py's headline one).
"""
import csv, json, os, re, subprocess, sys

rep, title, cmd = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes.sum.per_second']
print(f"# {title}\n\nCommand (under gpurun, after the same command had exited 0 without ncu):\n`{cmd}`\n")
print("ncu times are cold-cache and serialised; the live CUDA-event shares are in the bench JSON (`roofline.kernel_ms`).\n")
traffic = {}
for r in rows[2:]:
    short = re.sub(r'void <unnamed>::', '', r[h.index('Kernel Name')]).split('(')[0]
    print(f"## {short}\n\n| metric | value | unit |\n|---|---|---|")
    for k in want:
        if k in h:
            i = h.index(k)
            print(f"| {k} | {r[i]} | {units[i]} |")
    st = sorted(((float(r[i] or 0), x.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                 for i, x in enumerate(h) if 'issue_stalled' in x and 'per_issue_active' in x), reverse=True)
    print("\nStall cycles per issued instruction: " + ", ".join(f"{n} {v:.1f}" for v, n in st if v > 0.5) + "\n")
    traffic[short] = (float(r[h.index('dram__bytes_read.sum')]) + float(r[h.index('dram__bytes_write.sum')])) * 1e6
if "--no-traffic" in sys.argv[4:]:
    sys.exit(0)
# one entry per kernel of the captured timestep; the total is what bench.py reports as roofline.traffic
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump({"dram_bytes_per_step": int(sum(traffic.values())),
           "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of every kernel of ONE timestep "
                     "(tools/ncu_summary.py; " + os.path.basename(rep) + ")",
           "kernels": {k: int(v) for k, v in traffic.items()}},
          open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
