#!/usr/bin/env python
"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: ncu_launch_summary.py launches.csv"""
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    k = row["Kernel Name"].split("(")[0]; v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg[k][0] += 1; agg[k][1] += v; agg[k][2] = max(agg[k][2], v)
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold-cache: compare shares)")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:44s} n={v[0]:4d} total={v[1] / 1e3:9.2f} ms  avg={v[1] / v[0]:9.1f} us  max={v[2]:9.1f} us  share={v[1] / tot:.3f}")
