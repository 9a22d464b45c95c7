#!/usr/bin/env python
"""Prints the key figures of bench.py JSON lines.  Usage: show_bench.py file.json [...]"""
import json, sys
for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(f"{fn}: unreadable ({e})"); continue
    r = d.get("roofline", {}); c = d.get("counters", {})
    print(f"{fn}: {d['value']:.2f} {d['unit']}  e2e {d['e2e']['value']:.2f}  ms/step {d.get('ms_per_step') or 0:.1f}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
    if r: print(f"    roofline {r['kernel']}: {r['achieved']:.0f} GB/s = {r['frac']:.3f} of {r['peak_source']} {r['peak']:.0f}; share {({k: round(v, 3) for k, v in r['share_of_step'].items()})}")
    if d.get("phases_ms_per_step"): print("    phases/step", {k: (round(v, 1) if isinstance(v, (int, float)) else v) for k, v in d["phases_ms_per_step"].items()})
    if d.get("cpu_baseline"): print(f"    cpu {d['cpu_baseline']['value']:.3f} on {d['cpu_baseline']['cores']} cores")
    if c: print(f"    overflows {c.get('capacity_overflows')} segments {c.get('segments')} nodes {c.get('nodes_visited')} tris {c.get('tris_tested')}")
