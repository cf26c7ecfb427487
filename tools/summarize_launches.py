"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel totals and shares.
Usage: python tools/summarize_launches.py profiles/xyz.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if not hdr_i:
    raise SystemExit("no kernels in this file")
h, data = rows[hdr_i[0]], rows[hdr_i[0] + 1:]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[ki])[:100]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"launches {len(data)}  total {tot:.1f} us")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:12.1f} us {100 * t / tot:5.1f}%  n={c:5d} avg={t / c:9.1f}  {n}")
