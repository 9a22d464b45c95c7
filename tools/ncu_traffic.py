#!/usr/bin/env python
"""DRAM traffic per wavefront iteration of a phase's kernels from an ncu metrics pass (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv)
of tools/probe.py; merges the result into profiles/ncu_traffic.json (what bench.py reports as roofline.traffic).
Usage: ncu_traffic.py <launches.csv> <workload> <phase kernel name> <spp_per_step> <capture label> <kernel-regex of the phase> <iteration-marker kernel>"""
import collections, csv, json, os, re, sys
fn, workload, phase, spp, label, rx, marker = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5], re.compile(sys.argv[6]), sys.argv[7]
rows = list(csv.DictReader(l for l in open(fn) if not l.startswith("==")))
per = collections.defaultdict(lambda: collections.defaultdict(float)); iters = 0
ids = set()
for r in rows:
    k = r["Kernel Name"].split("(")[0]; m = r["Metric Name"]
    try: v = float(r["Metric Value"].replace(",", ""))
    except ValueError: continue
    u = r["Metric Unit"]
    if m.startswith("dram__bytes"): v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per[k][m] += v
    if marker in k and m == "gpu__time_duration.sum": iters += 1
tot = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for k, d in per.items() if rx.search(k))
print("iterations %d; phase kernels:" % iters)
for k, d in sorted(per.items(), key=lambda x: -(x[1]["dram__bytes_read.sum"] + x[1]["dram__bytes_write.sum"])):
    if rx.search(k): print("  %-32s read %8.1f MB  write %8.1f MB per iteration" % (k, d["dram__bytes_read.sum"] / 1e6 / max(1, iters), d["dram__bytes_write.sum"] / 1e6 / max(1, iters)))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
j = json.load(open(out)) if os.path.exists(out) else {}
j.setdefault(workload, {})[phase] = {"bytes": tot / max(1, iters), "capture": label, "spp_per_step": spp, "iterations": iters}
json.dump(j, open(out, "w"), indent=1)
print("traffic per iteration: %.1f MB -> %s" % (tot / max(1, iters) / 1e6, out))
