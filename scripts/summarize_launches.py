"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("tnpy::", "")[:64]
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches, {tot / 1e6:.3f} ms summed device time (cold-cache, serialised: compare shares)")
print(f"{'ms':>10} {'count':>6} {'share':>6} {'avg us':>9}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6:10.3f} {v[0]:6d} {100 * v[1] / tot:5.1f}% {v[1] / v[0] / 1e3:9.1f}  {k}")
