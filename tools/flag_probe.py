"""GPU probe: steady-state render time of the bench workload under different WTGPU_RENDER_* flags."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from wave_tracer_b200 import scenes, render, GpuScene
integ = sys.argv[1] if len(sys.argv) > 1 else "plt_bdpt"
b = scenes.double_slits(res=1440, spp=1024, integrator=integ, lut=(2048, 1024)).build()
gs = GpuScene(b, 0)
for flags in (0, 2, 8, 10, 0, 2):
    for rep in range(3):
        t = time.time(); _, _, st = render(b, spp=1024, sample_range=(rep * 16, rep * 16 + 16), gpu_scene=gs, flags=flags); dt = time.time() - t
    print("flags=%d: gpu %.1f ms wall %.1f ms -> %.2f Msamples/s; phases gen %.1f trav %.1f sort %.1f shade %.1f conn %.1f iters %d" % (
        flags, st["gpu_ms"], dt * 1e3, st["samples"] / st["gpu_ms"] / 1e3, st["generate_ms"], st["traverse_ms"], st["sort_ms"], st["shade_ms"], st["connect_ms"], st["iterations"]), flush=True)
