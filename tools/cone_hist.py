#!/usr/bin/env python
"""CPU diagnostic (oracle = test infrastructure): histogram of cone-query sizes of a bench workload.  python tools/cone_hist.py <workload> <res> <spp> [small]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, _oracle
wl, res, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]); small = len(sys.argv) > 4
sc, desc, tsz = bench.make_scene(wl, res, spp, small=small); b = sc.build(table_size=tsz)
L = _oracle.lib(); L.oracle_debug_cone_hist.argtypes = [C.c_int, C.POINTER(C.c_uint64)]
L.oracle_debug_cone_hist(1, None)
t = time.time(); _, _, st = _oracle.render(b, spp=spp); dt = time.time() - t
h = (C.c_uint64 * 96)(); L.oracle_debug_cone_hist(0, h)
print("tris %d; %d samples in %.1f s (%.4f Msamples/s, %d threads); segments %d cone casts %d tris tested %d" % (b.desc.n_tris, st["samples"], dt, st["samples"] / dt / 1e6, st["threads"], st["segments"], st["cone_casts"], st["tris"]))
print("bucket(<2^b)  queries_by_tested  tested_sum  share   | queries_by_accepted")
tot = sum(h[64 + i] for i in range(32)) or 1
for i in range(32):
    if h[i] or h[32 + i]: print("%2d %12d %14d %6.3f | %12d" % (i, h[i], h[64 + i], h[64 + i] / tot, h[32 + i]))
print("queries >= 512 accepted: %d; whose closest distance improved after the 512th accepted triangle: %d (+%d further improvements)" % (sum(h[32 + i] for i in range(10, 32)), h[94], h[95]))
