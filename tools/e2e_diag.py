#!/usr/bin/env python
"""Where the end-to-end time of a host-buffer render goes: scene upload (wtgpu_scene_create), wtgpu_render with host films, scene destroy.
Usage: e2e_diag.py [etoile|double_slits|cornell]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from wave_tracer_b200 import scenes, GpuScene
wl = sys.argv[1] if len(sys.argv) > 1 else "etoile"
if wl == "etoile": b = scenes.etoile_like(res=720, spp=1024).build(); S = 16
elif wl == "cornell": b = scenes.cornell_like(res=1440, spp=1024, integrator="plt_bdpt", fsd=True, lut=(2048, 1024)).build(); S = 2
else: b = scenes.double_slits(res=1440, spp=1024, integrator="plt_bdpt", lut=(2048, 1024)).build(); S = 16
W, H, Cn = b.width, b.height, b.channels
for it in range(4):
    t0 = time.time(); gs = GpuScene(b, 0); t1 = time.time()
    block = np.zeros((H, W, Cn, 2), np.float32); light = np.zeros((H, W, Cn), np.float32); t2 = time.time()
    st = gs.render_into(block.ctypes.data_as(C.c_void_p), light.ctypes.data_as(C.c_void_p), 1024, 0x5EED, (it * S, it * S + S), None, False, 0, 0, None, True); t3 = time.time()
    gs.close(); t4 = time.time()
    print(f"{wl} it{it}: scene_create {1e3 * (t1 - t0):7.1f} ms  host film alloc {1e3 * (t2 - t1):6.1f}  wtgpu_render {1e3 * (t3 - t2):7.1f} (gpu_ms {st['gpu_ms']:.1f})  destroy {1e3 * (t4 - t3):6.1f}")
