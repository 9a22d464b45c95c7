"""GPU probe: plt_bdpt throughput on the double_slits workload (development aid, not a bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from wave_tracer_b200 import scenes, render, GpuScene
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1440
scene = sys.argv[2] if len(sys.argv) > 2 else "ds"
for fsd in (True, False):
    if scene == "ds":
        b = scenes.double_slits(res=res, spp=1024, integrator="plt_bdpt", fsd=fsd, lut=(2048, 1024)).build()
    else:
        if fsd: continue
        b = scenes.cornell_like(res=res, spp=1024, integrator="plt_bdpt").build()
    gs = GpuScene(b, 0)
    for spp, pool, flags in ((4, 0, 2), (4, 1 << 19, 2), (4, 0, 10)):
        t = time.time(); _, _, st = render(b, spp=1024, sample_range=(0, spp), gpu_scene=gs, pool_size=pool, flags=flags); dt = time.time() - t
        print("gpu %s fsd=%s spp=%d pool=%d flags=%d: %.1f ms, %.3f Msamples/s (wall %.1f ms) iters=%d gen %.1f trav %.1f sort %.1f shade+conn %.1f ovf=%d verts=%d conn=%d" % (
            scene, fsd, spp, pool, flags, st["gpu_ms"], st["samples"] / st["gpu_ms"] / 1e3, dt * 1e3, st["iterations"], st["generate_ms"], st["traverse_ms"], st["sort_ms"], st["shade_ms"],
            st["capacity_overflows"], st["segments"], st["shaded_paths"]), flush=True)
