#!/usr/bin/env python
"""Executed warp instructions and stall samples of a kernel per source line / inline chain.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv <object-with-lineinfo (.so/.o)> <mangled-section-substring> [top] [depth]

Joins the per-instruction columns of the ncu source page with the inlined file:line chain printed by
`nvdisasm --print-line-info-inline` (carried forward between line markers).
"""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, obj, sec = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 3
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
chain = {}
for cubin in [f for f in os.listdir(tmp) if f.endswith(".cubin")]:
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    in_sec = False
    pend, last = [], ()
    for line in dis.splitlines():
        if line.startswith("//--------------------- .text."):
            in_sec = sec in line
            pend, last = [], ()
            continue
        if not in_sec:
            continue
        m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", line)
        if m:
            pend.append(f"{os.path.basename(m.group(1)).replace('wsi_', '').replace('.cuh', '')}:{m.group(2)}")
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pend:
                last = tuple(pend)
            chain[int(m.group(1), 16)] = last
            pend = []
rows = csv.reader(open(src_csv))
hdr = None
base = None
inst = collections.Counter(); smp = collections.Counter(); loc = collections.Counter()
inst_o = collections.Counter(); smp_o = collections.Counter()
n_all = s_all = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        iN, iS, iP = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < 6 or r[0] == "Kernel Name":
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    n, s = int(r[iN] or 0), int(r[iP] or 0)
    n_all += n; s_all += s
    ch = chain.get(addr - base, ())
    key = " <- ".join(ch[:depth])
    inst[key] += n; smp[key] += s
    ok = ch[-1] if ch else "?"
    if len(ch) >= 2:
        ok = ch[-2] + " <- " + ch[-1]
    inst_o[ok] += n; smp_o[ok] += s
    op = r[iS].split()
    op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
    if op.startswith("STL") or op.startswith("LDL"):
        loc[key] += n
print(f"{n_all} warp instructions, {s_all} samples")
print("--- by outer context (instructions, samples)")
for k, n in inst_o.most_common(top):
    print(f"{n:10d} {100*n/max(n_all,1):5.1f}%  smp {smp_o[k]:6d} {100*smp_o[k]/max(s_all,1):5.1f}%  {k}")
print("--- by inline chain (samples)")
for k, s in smp.most_common(top):
    print(f"{inst[k]:10d} {100*inst[k]/max(n_all,1):5.1f}%  smp {s:6d} {100*s/max(s_all,1):5.1f}%  local {loc[k]:8d}  {k}")
