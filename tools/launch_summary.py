#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, idi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
    per.setdefault((r[idi], r[ki][:64]), {})[r[mi]] = v * scale
agg = collections.OrderedDict()
for (_, name), m in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"{'ms total':>10s} {'n':>4s} {'share':>6s} {'ms/launch':>10s} {'GB/launch':>10s} {'GB/s':>8s}  kernel")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = a[2] / (a[1] * 1e-3) if a[1] > 0 and a[2] > 0 else 0
    print(f"{a[1]:10.2f} {a[0]:4d} {100 * a[1] / tot:5.1f}% {a[1] / a[0]:10.3f} {a[2] / a[0]:10.2f} {gbs:8.0f}  {n}")
print(f"{tot:10.2f} total")
