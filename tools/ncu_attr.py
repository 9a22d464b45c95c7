#!/usr/bin/env python
"""Attributes an `ncu --page source --csv` export (SASS view, one row per instruction) to the SOURCE FUNCTIONS of wave_tracer_b200/csrc:
stall samples, executed warp-instructions, lanes per instruction and static SASS size per function.  The SASS<->line map comes from the
in-tree library (built with -lineinfo) through cuobjdump/nvdisasm, so the library must be the build that was profiled.

Usage: ncu_attr.py <source.csv> <kernel-substring> [lib.so]        e.g.  ncu_attr.py gpurun_out/full_q_et_gtraverse_source.csv k_gtraverse
"""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wave_tracer_b200", "csrc")


def function_ranges():
    out = {}
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:WT_D|WT_DN|WT_NI|__global__|__device__|static|inline)\b[^;{]*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    for fn in os.listdir(CSRC):
        if not fn.endswith((".cuh", ".cu")): continue
        defs = []
        prev_template = False
        for i, line in enumerate(open(os.path.join(CSRC, fn)), 1):
            if line.startswith((" ", "\t", "//", "#", "}")):
                continue
            m = pat.match(line)
            if m and m.group(1) not in ("if", "for", "while", "switch", "return", "__launch_bounds__"):
                defs.append((i, m.group(1)))
            elif "__global__" in line:
                m2 = re.search(r"\)\s*([A-Za-z_][A-Za-z_0-9]*)\s*\(", line)
                if m2: defs.append((i, m2.group(1)))
        out[fn] = defs
    return out


def sass_lines(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    cubin = max(cubins, key=lambda f: os.path.getsize(os.path.join(tmp, f)))
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, take, cur = [], False, None
    for l in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):", l)
        if m:
            take = kernel_sub in m.group(1); cur = None
            continue
        if not take: continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm: cur = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
    return lines


def main():
    src, ksub = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "wave_tracer_b200", "libwt_b200.so")
    rows = list(csv.reader(open(src)))
    hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
    m = sass_lines(lib, ksub)
    if len(m) != len(data):
        print(f"warning: library has {len(m)} SASS instructions for *{ksub}*, the capture {len(data)}: attribution skipped (profile a matching build)"); return
    fr = function_ranges()
    def fn_of(k):
        if k is None: return "?"
        f, l = k
        name = "?"
        for i, n in fr.get(f, []):
            if i <= l: name = n
            else: break
        return f"{f}:{name}" if f in fr else f
    num = lambda r, n: float(r[col[n]]) if r[col[n]] not in ("", "-") else 0.0
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for r, k in zip(data, m):
        a = agg[fn_of(k)]
        a[0] += num(r, "# Samples"); a[1] += num(r, "Instructions Executed"); a[2] += num(r, "Thread Instructions Executed"); a[3] += 1
    ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
    print(f"{rows[0][1]}: {len(data)} SASS instructions ({len(data) * 16 // 1024} KB), {ti:.3g} warp-instructions, {ts:.0f} stall samples")
    print(f"{'function':44s} {'SASS':>6s} {'samples%':>9s} {'inst%':>7s} {'lanes':>6s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:32]:
        print(f"{k:44s} {v[3]:6d} {100 * v[0] / ts:9.1f} {100 * v[1] / ti:7.1f} {v[2] / max(v[1], 1):6.1f}")


if __name__ == "__main__":
    main()
