#!/usr/bin/env python
"""Correlate an ncu SASS profile with source lines (no GUI needed).

    ncu -i prof.ncu-rep --page source --csv -k regex:<kernel> > sass.csv
    python tools/ncu_hotspots.py sass.csv wsi_b200/csrc/libwsi_b200.so <kernel-substring> [top] [mangled-section-substring]

Uses `cuobjdump -xelf` + `nvdisasm --print-line-info-inline` for the offset -> file:line map and
reports executed warp instructions and stall samples per source line and per device function.
"""
import collections, csv, os, re, subprocess, sys, tempfile

sass_csv, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sec_name = sys.argv[5] if len(sys.argv) > 5 else kname  # substring of the (mangled) .text section
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

# offset -> (file:line innermost, function label) for the wanted kernel's .text section
loc, func = {}, {}
in_sec = False
pend = None
fresh = True
cur_func = None
for line in dis.splitlines():
    if line.startswith("//--------------------- .text."):
        in_sec = sec_name in line
        pend = None
        cur_func = "kernel"
        continue
    if not in_sec:
        continue
    m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", line)
    if m:
        if fresh:
            pend = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            fresh = False
        continue
    m = re.match(r"^(\$?[A-Za-z_][\w$.]*):\s*$", line)
    if m and not m.group(1).startswith(".L"):
        cur_func = m.group(1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        off = int(m.group(1), 16)
        loc[off] = pend or "?"
        func[off] = cur_func
        fresh = True

rows = csv.reader(open(sass_csv))
hdr = None
base = None
by_line = collections.Counter(); smp_line = collections.Counter()
by_func = collections.Counter(); smp_func = collections.Counter()
tot = ts = 0
seen_kernel = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if seen_kernel and base is not None:
            break
        seen_kernel = kname in r[1]
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        iN, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if not seen_kernel or hdr is None or len(r) < 6:
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    off = addr - base
    n, s = int(r[iN] or 0), int(r[iS] or 0)
    tot += n; ts += s
    by_line[loc.get(off, "?")] += n; smp_line[loc.get(off, "?")] += s
    by_func[func.get(off, "?")] += n; smp_func[func.get(off, "?")] += s
print(f"kernel {kname}: {tot} warp instructions, {ts} samples")
print("--- per device function")
for f, n in by_func.most_common(25):
    short = f.split("$")[-1]
    try:
        short = subprocess.run(["c++filt", short], capture_output=True, text=True).stdout.strip()[:150]
    except Exception:
        pass
    print(f"{n:10d} {100*n/max(tot,1):5.1f}%  samples {smp_func[f]:6d} {100*smp_func[f]/max(ts,1):5.1f}%  {short}")
print("--- per source line (by samples)")
for l, s in smp_line.most_common(top):
    print(f"{by_line[l]:10d} {100*by_line[l]/max(tot,1):5.1f}%  samples {s:6d} {100*s/max(ts,1):5.1f}%  {l}")
