#!/usr/bin/env python
"""Per-kernel time of the LAST step in an ncu launch list (gpu__time_duration.sum):
   python tools/launch_shares.py gpurun_out/x_launches.csv LAUNCHES_PER_STEP [out.txt]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [(re.sub(r"\(.*", "", r[ki]).replace("syn::", "").replace("void ", ""), float(r[vi].replace(",", ""))) for r in rows[1:]]
per = int(sys.argv[2])
last = names[-per:]
agg = collections.OrderedDict()
for (n, v) in last:
    agg.setdefault(n, []).append(v)
tot = sum(v for (_, v) in last)
out = [f"{sys.argv[1]}: last {per} launches (one step), ncu --metrics gpu__time_duration.sum --clock-control none"]
for (k, v) in agg.items():
    out.append(f"{k:56s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:8.1f}us total={sum(v) / 1e3:8.1f}us {100 * sum(v) / tot:5.1f}%")
out.append(f"sum of kernel time of the step (serialised, cold cache): {tot / 1e3:.1f} us")
print("\n".join(out))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(out) + "\n")
