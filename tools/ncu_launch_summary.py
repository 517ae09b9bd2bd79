#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / mean / total (us)."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0]); a[0] += 1; a[1] += us; a[2] = min(a[2], us); a[3] = max(a[3], us)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'mean_us':>9s} {'min_us':>9s} {'max_us':>9s} {'total_us':>10s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {a[0]:5d} {a[1]/a[0]:9.1f} {a[2]:9.1f} {a[3]:9.1f} {a[1]:10.1f} {100*a[1]/tot:5.1f}%")
