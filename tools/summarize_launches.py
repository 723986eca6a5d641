#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (times are cold-cache and serialised:
compare SHARES, not absolutes).   python tools/summarize_launches.py gpurun_out/launches_r1.csv > profiles/launches_r1.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"# ncu launch list summary: {path}\n")
print(f"{sum(cnt.values())} launches, {T:.1f} ms total under ncu (serialised, cold caches)\n")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"| `{k}` | {cnt[k]} | {v:.2f} | {100 * v / T:.1f}% |")
