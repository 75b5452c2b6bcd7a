#!/usr/bin/env python3
"""Launch list summary from `ncu --metrics gpu__time_duration.sum --csv --log-file F`:
    python tools/launch_list.py gpurun_out/launches_ss_final.csv > profiles/r02_launches_steady_state.txt"""
import collections
import csv
import re
import sys

rows = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
agg = collections.OrderedDict()
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name)
    us = float(r["Metric Value"].replace(",", "")) / (1000.0 if r["Metric Unit"] == "ns" else 1.0)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'mean us':>9s} {'max us':>9s} {'total ms':>10s}  share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {a[0]:8d} {a[1] / a[0]:9.1f} {a[2]:9.1f} {a[1] / 1000:10.2f} {100 * a[1] / tot:5.1f}%")
