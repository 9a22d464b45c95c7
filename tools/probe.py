#!/usr/bin/env python
"""One render of a bench workload at a chosen size with every statistic printed (GPU box): python tools/probe.py <workload> <res> <spp> [flags] [pool]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from wave_tracer_b200 import GpuScene, render
wl, res, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 2
pool = int(sys.argv[5]) if len(sys.argv) > 5 else 0
t = time.time(); sc, desc, tsz = bench.make_scene(wl, res, spp); b = sc.build(table_size=tsz); print("build %.1f s; tris %d edges %d nodes %d" % (time.time() - t, b.desc.n_tris, b.desc.n_edges, b.desc.n_nodes), flush=True)
gs = GpuScene(b, 0)
if os.environ.get("PROBE_CAPS"): gs.set_capacities([int(x) for x in os.environ["PROBE_CAPS"].split(",")])
for rep in range(int(os.environ.get("PROBE_REPS", "3"))):
    t = time.time(); _, _, st = render(b, spp=spp, gpu_scene=gs, flags=flags, pool_size=pool); dt = time.time() - t
    print("render %d: %.2f s wall, %.1f ms gpu, %.3f Msamples/s; passes %d pool %d caps %s; iterations %d launches %d" % (rep, dt, st["gpu_ms"], st["samples"] / st["gpu_ms"] / 1e3, st["passes"], st["pool_used"], gs.capacities(), st["iterations"], st["kernel_launches"]), flush=True)
    print("   phases ms: gen %.1f trav %.1f sort %.1f shade %.1f conn %.1f | segments %d nodes/seg %.1f tris/seg %.1f shadow %d" % (st["generate_ms"], st["traverse_ms"], st["sort_ms"], st["shade_ms"], st["connect_ms"], st["segments"], st["nodes_visited"] / max(1, st["segments"]), st["tris_tested"] / max(1, st["segments"]), st["shadow_casts"]), flush=True)
