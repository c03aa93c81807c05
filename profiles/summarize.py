#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches.csv          > profiles/r01_launches.txt
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep          > profiles/r01_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def launches(path):
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        agg.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"].replace(",", "")))
    total = sum(sum(v) / len(v) for v in agg.values())
    print(f"{'kernel':70s} {'launches':>8s} {'mean us':>10s} {'share':>7s}")
    for k, v in agg.items():
        mean = sum(v) / len(v)
        print(f"{k[:70]:70s} {len(v):8d} {mean / 1e3:10.1f} {100 * mean / total:6.1f}%")
    print(f"{'sum of per-launch means (one pass)':70s} {'':8s} {total / 1e3:10.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("-----", r[idx["Kernel Name"]])
        for m in FULL_METRICS:
            if m in idx:
                print(f"  {m:70s} {r[idx[m]]:>22s} {units[idx[m]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
