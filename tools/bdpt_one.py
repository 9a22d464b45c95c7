"""One plt_bdpt render (profiling target): bdpt_one.py <scene ds|cb> <res> <fsd 0|1> <spp>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from wave_tracer_b200 import scenes, render, GpuScene
scene = sys.argv[1]; res = int(sys.argv[2]); fsd = sys.argv[3] != "0"; spp = int(sys.argv[4])
if scene == "ds":
    b = scenes.double_slits(res=res, spp=1024, integrator="plt_bdpt", fsd=fsd, lut=(2048, 1024)).build()
else:
    b = scenes.cornell_like(res=res, spp=1024, integrator="plt_bdpt").build()
gs = GpuScene(b, 0)
_, _, st = render(b, spp=1024, sample_range=(0, spp), gpu_scene=gs)
print(st["gpu_ms"], st["samples"])
