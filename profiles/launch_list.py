#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean / min duration per kernel and
each kernel's share of the listed time.

    python profiles/launch_list.py gpurun_out/r2_shard8.csv [launches_per_pass_divisor] > profiles/r02_launches_....txt
"""
import collections
import csv
import sys

path = sys.argv[1]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        agg.setdefault(d["Kernel Name"][:78], []).append(v)
total = sum(sum(v) / len(v) * (len(v) / div) for v in agg.values())
print(f"{'kernel':78s} {'launches':>8s} {'per pass':>8s} {'mean us':>9s} {'min us':>8s} {'share':>6s}")
for k, v in agg.items():
    m = sum(v) / len(v)
    print(f"{k:78s} {len(v):8d} {len(v) / div:8.2f} {m:9.1f} {min(v):8.1f} {100 * m * (len(v) / div) / total:5.1f}%")
print(f"{'sum of (mean x launches per pass)':78s} {'':8s} {'':8s} {total:9.1f}")
