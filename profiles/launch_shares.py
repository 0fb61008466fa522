#!/usr/bin/env python
"""Shares of a step by kernel from an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv).  usage: python profiles/launch_shares.py gpurun_out/launches.csv "<command that was profiled>"
Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in rows[1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    name = r[ix["Kernel Name"]].split("(")[0][-70:]
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit, metric = r[ix["Metric Unit"]], r[ix["Metric Name"]]
    if metric == "gpu__time_duration.sum":
        acc[name][0] += 1
        acc[name][1] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    else:
        mb = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        acc[name][2 if metric == "dram__bytes_read.sum" else 3] += mb
total = sum(a[1] for a in acc.values())
print("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none:", sys.argv[2] if len(sys.argv) > 2 else "")
print("(cold-cache, serialised launches: compare shares, not absolutes; dram bytes are sums over the launches listed)")
for name, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:4d} launches {a[1]:12.3f} ms {100 * a[1] / total:5.1f}%  dram rd {a[2]:9.1f} MB wr {a[3]:9.1f} MB  {name}")
