#!/usr/bin/env python
"""Summarises an `ncu --page source --csv` export (SASS view): opcode histogram weighted by executed warp-instructions and stall samples,
and the hottest contiguous SASS regions.  Usage: ncu_source_summary.py file.csv [region_size]"""
import csv, sys, collections
fn = sys.argv[1]; R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rows = list(csv.reader(open(fn)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def num(r, name):
    try: return float(r[col[name]])
    except Exception: return 0.0
tot_i = sum(num(r, "Instructions Executed") for r in data); tot_s = sum(num(r, "# Samples") for r in data)
print(f"kernel {rows[0][1]}: {len(data)} SASS instructions, {tot_i:.3g} warp-instructions executed, {tot_s:.0f} stall samples")
ops = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in data:
    t = r[col["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op][0] += num(r, "Instructions Executed"); ops[op][1] += num(r, "# Samples"); ops[op][2] += num(r, "Thread Instructions Executed")
print("opcode       inst%  samples%  avg-threads")
for op, v in sorted(ops.items(), key=lambda x: -x[1][1])[:22]:
    print(f"  {op:10s} {100 * v[0] / tot_i:6.2f} {100 * v[1] / max(tot_s, 1):8.2f} {v[2] / max(v[0], 1):8.1f}")
print(f"hottest {R}-instruction regions (by stall samples):")
regs = []
for i in range(0, len(data), R):
    blk = data[i:i + R]
    regs.append((sum(num(r, "# Samples") for r in blk), sum(num(r, "Instructions Executed") for r in blk), i, blk))
for s, n, i, blk in sorted(regs, key=lambda x: -x[0])[:8]:
    stalls = collections.Counter()
    for r in blk:
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h: stalls[h] += num(r, h)
    top = ", ".join(f"{k[6:]}={v:.0f}" for k, v in stalls.most_common(3))
    thr = sum(num(r, "Thread Instructions Executed") for r in blk) / max(n, 1)
    print(f"  [{i:5d}..{i + R:5d}) samples {100 * s / max(tot_s, 1):5.1f}%  inst {100 * n / tot_i:5.1f}%  avg-threads {thr:4.1f}  stalls: {top}   e.g. {blk[len(blk)//2][col['Source']].strip()[:50]}")
