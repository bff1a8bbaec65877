"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and launches per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += v * scale
tot = sum(v for _, v in agg.values())
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:10.3f} ms {100 * v / tot:5.1f}% {c:5d}  {k}")
print(f"{tot:10.3f} ms total")
