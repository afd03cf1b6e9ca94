#!/usr/bin/env python
"""List the fp64 division call sites that take the slow path (__cuda_sm20_div_rn_f64_full) in an
ncu source-page CSV, with executed counts and the source line (inline chain) of each site.

    ncu -i prof.ncu-rep --page source --csv -k regex:<kernel> > sass.csv
    python tools/ncu_slowdiv.py sass.csv wsi_b200/csrc/libwsi_b200.so <mangled-section-substring>
"""
import csv, os, re, subprocess, sys, tempfile

sass_csv, so, sec = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
loc, func, labels = {}, {}, {}
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
            chain = []
            fresh = False
        chain.append(os.path.basename(m.group(1)) + ":" + m.group(2))
        continue
    m = re.match(r"^(\$?[A-Za-z_][\w$.]*):\s*$", line)
    if m and not m.group(1).startswith(".L"):
        cur = m.group(1)
        labels.setdefault(cur, None)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        off = int(m.group(1), 16)
        loc[off] = " <- ".join(chain[:4])
        func[off] = cur
        if labels.get(cur) is None:
            labels[cur] = off
        fresh = True
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iS, iN, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = int(rows[2][0], 16)
full = [k for k in labels if k and "div_rn_f64_full" in k]
target = hex(base + labels[full[0]])[2:]
out = {}
for r in rows[2:]:
    if len(r) <= iT or "CALL" not in r[iS] or target not in r[iS] or int(r[iN]) == 0:
        continue
    off = int(r[0], 16) - base
    key = ((func.get(off) or "")[-40:], loc.get(off))
    a = out.setdefault(key, [0, 0, 0])
    a[0] += int(r[iN]); a[1] += int(r[iT]); a[2] += 1
for k, v in sorted(out.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{v[0]:8d} warp-calls {v[1]:9d} threads  sites {v[2]:2d}  {k[0]}  {k[1]}")
