#!/bin/bash
# gpurun job: device time of every launch of a few config-2 timesteps (cold-cache, serialised: compare shares)
tag=${1:-launches}
mkdir -p gpurun_out
export WSI_B200_NO_GRAPH=1
python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --secondary "" > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 12 -c 16 --csv --log-file gpurun_out/${tag}.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --secondary "" > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}.csv")) if len(r) > 10]
h = rows[0]; iK, iM, iV = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iID = h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[iID], r[iK][:60]), {})[r[iM]] = float(r[iV].replace(",", ""))
for (i, k), m in list(per.items())[:10]:
    print(f"{k:60s} {m.get('gpu__time_duration.sum',0)/1e3:8.1f} us  rd {m.get('dram__bytes_read.sum',0)/1e6:7.1f} MB  wr {m.get('dram__bytes_write.sum',0)/1e6:7.1f} MB  inst {m.get('smsp__inst_executed.sum',0)/1e6:6.1f} M")
PY
