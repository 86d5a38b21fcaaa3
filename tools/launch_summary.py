#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/launch_summary.py launches.csv [top]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
    agg[re.sub(r"\(.*", "", row["Kernel Name"])[:72]].append(v)
tot = sum(sum(v) for v in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
print(f"{'kernel':74s} {'n':>6s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
    print(f"{k:74s} {len(v):6d} {sum(v) / len(v):9.2f} {sum(v) / tot:7.1%}")
