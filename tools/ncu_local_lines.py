#!/usr/bin/env python
"""Executed local-memory instructions (spills / stack) of a kernel per source line.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_local_lines.py src.csv <object-with-lineinfo (.so/.o)> <mangled-section-substring> [top]

Joins the per-instruction "Instructions Executed" of the ncu source page with the inlined
file:line chain from nvdisasm, and prints where the executed STL/LDL instructions come from.
"""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, obj, sec = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
chain = {}
for cubin in [f for f in os.listdir(tmp) if f.endswith(".cubin")]:
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    in_sec = False
    pend = []
    for line in dis.splitlines():
        if line.startswith("//--------------------- .text."):
            in_sec = sec in line
            pend = []
            last = ()
            continue
        if not in_sec:
            continue
        m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", line)
        if m:
            pend.append(f"{os.path.basename(m.group(1))}:{m.group(2)}")
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
tot = collections.Counter()
by_inner = collections.Counter()
by_outer = collections.Counter()
by_chain = collections.Counter()
n_all = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        iN, iS = hdr.index("Instructions Executed"), hdr.index("Source")
        continue
    if hdr is None or len(r) < 6 or r[0] == "Kernel Name":
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    n = int(r[iN] or 0)
    n_all += n
    op = r[iS].split()[0] if r[iS] else ""
    if r[iS].startswith("@"):
        op = r[iS].split()[1]
    kind = "STL" if op.startswith("STL") else "LDL" if op.startswith("LDL") else None
    if not kind:
        continue
    ch = chain.get(addr - base, ())
    tot[kind] += n
    by_inner[(kind, ch[0] if ch else "?")] += n
    by_outer[(kind, ch[-1] if ch else "?")] += n
    by_chain[(kind, " <- ".join(ch[:4]))] += n
print(f"{n_all} warp instructions; local: {dict(tot)}")
for title, c in (("innermost line", by_inner), ("outermost line", by_outer), ("inline chain", by_chain)):
    print(f"--- by {title}")
    for (k, l), n in c.most_common(top):
        print(f"{n:10d} {k} {l}")
