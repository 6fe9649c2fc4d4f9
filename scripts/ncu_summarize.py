#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import collections, csv, statistics, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
d = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    d.setdefault(name, []).append(v)
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':48s} {'launches':>8s} {'median_us':>10s} {'mean_us':>9s} {'max_us':>8s} {'total_us':>10s} {'share':>6s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:48]:48s} {len(v):8d} {statistics.median(v):10.2f} {sum(v)/len(v):9.2f} {max(v):8.2f} {sum(v):10.1f} {100*sum(v)/tot:5.1f}%")
