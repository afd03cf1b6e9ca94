#!/usr/bin/env python
"""Executed warp instructions and stall samples of a kernel aggregated by the OUTERMOST inlined
call site inside each physical function (ncu source-page CSV + nvdisasm inline line info).

    python tools/ncu_callsites.py sass.csv lib.so <mangled-section-substring> [top]
"""
import collections, csv, os, re, subprocess, sys, tempfile

sass_csv, so, sec = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
site, func = {}, {}
insec, chain, fresh, cur = False, [], True, None
for line in dis.splitlines():
    if line.startswith("//--------------------- .text."):
        insec = sec in line
        continue
    if not insec:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append(os.path.basename(m.group(1)) + ":" + m.group(2))
        continue
    m = re.match(r"^(\$?[A-Za-z_][\w$.]*):\s*$", line)
    if m and not m.group(1).startswith(".L"):
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        off = int(m.group(1), 16)
        site[off] = chain[-1] if chain else "?"
        func[off] = cur
        fresh = True
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iN, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][0], 16)
agg = collections.Counter(); smp = collections.Counter()
tot = ts = 0
for r in rows[2:]:
    if len(r) <= max(iN, iS):
        continue
    off = int(r[0], 16) - base
    f = (func.get(off) or "kernel").split("$")[-1]
    f = re.sub(r"^_ZN3wsi\d+", "", f)[:28]
    key = (f, site.get(off, "?"))
    n, s_ = int(r[iN] or 0), int(r[iS] or 0)
    agg[key] += n; smp[key] += s_; tot += n; ts += s_
print(f"{tot} warp instructions, {ts} samples")
for k, n in agg.most_common(top):
    print(f"{n:10d} {100*n/tot:5.1f}%  samples {100*smp[k]/max(ts,1):5.1f}%  {k[0]:30s} {k[1]}")
