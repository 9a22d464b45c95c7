"""Render driver: the Python face of wtgpu_render (include/wtgpu.h), replacing scene_renderer_t::render
(/root/reference/src/scene/render.cpp:381-579) for the sensor of a built scene.

Multi-GPU: every rank renders a disjoint sample range of every element into its own full-size film
(the reference's per-thread light-image scheme, film_storage.hpp:155-158,276-288) and the films are summed with ONE
NCCL reduce at the end (torch.distributed).  RNG streams are keyed by (pixel, sample), so the image does not depend
on the number of GPUs up to f32 summation order.
"""
import ctypes as C
import numpy as np

from . import _abi as A


class GpuScene:
    """Device-resident scene (wtgpu_scene).  Fails loudly without the native library or a GPU."""
    def __init__(self, built, device=0):
        self.built, self.device = built, device
        self.handle = C.c_void_p()
        A.check(A.lib().wtgpu_scene_create(C.byref(built.desc), device, C.byref(self.handle)), "wtgpu_scene_create")

    def close(self):
        if self.handle:
            A.lib().wtgpu_scene_destroy(self.handle); self.handle = C.c_void_p()

    def __del__(self):
        try: self.close()
        except Exception: pass

    def capacities(self):
        """Row lengths of the per-path lists {cone triangles, edges, aperture segments, apertures per subpath, vertices per subpath} (include/wtgpu.h)."""
        out = (C.c_uint32 * 5)(); A.check(A.lib().wtgpu_get_capacities(self.handle, out), "wtgpu_get_capacities"); return list(out)

    def set_capacities(self, caps):
        A.check(A.lib().wtgpu_set_capacities(self.handle, (C.c_uint32 * 5)(*caps)), "wtgpu_set_capacities")

    def render_into(self, block_ptr, light_ptr, spp, seed=0x5EED, sample_range=None, tile=None, on_device=False, pool_size=0, flags=0, stream=None):
        b = self.built
        o = A.RenderOpts()
        o.seed, o.spp = seed, spp
        o.sample_begin, o.sample_end = sample_range if sample_range else (0, spp)
        o.tile_x0, o.tile_y0, o.tile_x1, o.tile_y1 = tile if tile else (0, 0, b.width, b.height)
        o.device, o.film_on_device, o.pool_size, o.flags = self.device, int(on_device), pool_size, flags
        o.stream = stream
        o.sampler = getattr(b, "sampler", 0)
        st = A.Stats()
        A.check(A.lib().wtgpu_render(self.handle, C.byref(o), block_ptr, light_ptr, C.byref(st)), "wtgpu_render")
        return st.as_dict()


def render(built, spp=None, seed=0x5EED, device=0, sample_range=None, tile=None, pool_size=0, flags=0, gpu_scene=None):
    """Host-buffer render through the C-ABI (host<->device copies inside the call).  Returns (film_block, film_light, stats)."""
    spp = spp or built.spp
    gs = gpu_scene or GpuScene(built, device)
    W, H, Cn = built.width, built.height, built.channels
    block = np.zeros((H, W, Cn, 2), np.float32); light = np.zeros((H, W, Cn), np.float32)
    st = gs.render_into(block.ctypes.data_as(C.c_void_p), light.ctypes.data_as(C.c_void_p), spp, seed, sample_range, tile, False, pool_size, flags, None)
    if gpu_scene is None:
        gs.close()
    return block, light, st


def develop(built, spp, film_block, film_light):
    """value/weight + light/spp (film_storage.hpp:256-291, 354-358)."""
    fb = np.asarray(film_block, np.float64); fl = np.asarray(film_light, np.float64)
    w = fb[..., 1]
    img = np.where(w > 0, fb[..., 0] / np.where(w > 0, w, 1), 0.0)
    return img + fl * (1.0 / spp if spp > 0 else 0.0)
