"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

from wave_tracer_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")                 # elementary functions from pmath.h (same bits as the device): GPU parity, CPU baseline
LIB_GLIBC = os.path.join(ORACLE_DIR, "liboracle_glibc.so")     # host libm as the reference calls it: the pins against the reference's own code


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "segments", "surface", "fsd", "null_", "splats", "nodes", "tris", "ray_casts", "cone_casts", "shadow_casts")] + \
               [("seconds", C.c_double), ("threads", C.c_uint32), ("pad_", C.c_uint32)]


_lib = None
_lib_glibc = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def lib_glibc():
    global _lib_glibc
    if _lib_glibc is None:
        if not os.path.exists(LIB_GLIBC):
            build()
        _lib_glibc = _bind(C.CDLL(LIB_GLIBC))
    return _lib_glibc


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = _bind(C.CDLL(LIB))
    return _lib


def _bind(L):
    if True:
        if True:
            pass
        L.oracle_render.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.RenderOpts), C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(OracleStats)]
        L.oracle_intersect_rays.argtypes = [C.POINTER(A.SceneDesc), C.c_uint32, C.POINTER(A.RayQuery), C.POINTER(A.RayHit)]
        L.oracle_intersect_rays_bruteforce.argtypes = L.oracle_intersect_rays.argtypes
        L.oracle_shadow_rays.argtypes = [C.POINTER(A.SceneDesc), C.c_uint32, C.POINTER(A.RayQuery), C.POINTER(C.c_uint32)]
        L.oracle_intersect_cones.argtypes = [C.POINTER(A.SceneDesc), C.c_uint32, C.POINTER(A.ConeQuery), C.POINTER(A.ConeHit)]
        L.oracle_cone_closest_bruteforce.argtypes = [C.POINTER(A.SceneDesc), C.c_uint32, C.POINTER(A.ConeQuery), C.POINTER(C.c_float)]
        L.oracle_rng.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        L.oracle_sobol_batch.argtypes = [C.POINTER(A.SobolEntry), C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.oracle_sobol_matrices.argtypes = [C.POINTER(A.SobolEntry), C.POINTER(C.c_int32)]
        L.oracle_svd.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_utdf.argtypes = [C.c_float, C.POINTER(C.c_float)]
        L.oracle_cerfc_rot45.argtypes = [C.c_double, C.POINTER(C.c_double)]
        L.oracle_fresnel.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_mub_sbp.argtypes = [C.c_float, C.c_float]; L.oracle_mub_sbp.restype = C.c_float
        L.oracle_bsdf_albedo.argtypes = [C.POINTER(A.SceneDesc), C.c_int32, C.POINTER(C.c_float), C.c_float, C.c_uint32, C.c_uint64]
        L.oracle_bsdf_albedo.restype = C.c_float
        L.oracle_profile_eval.argtypes = [C.POINTER(A.SceneDesc), C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
        L.oracle_profile_check.argtypes = [C.POINTER(A.SceneDesc), C.c_int32, C.POINTER(C.c_float), C.c_float, C.c_uint32, C.c_uint64, C.POINTER(C.c_float)]
        L.oracle_fuzz_cone_quick_reject.argtypes = [C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_fuzz_ray_cull.argtypes = [C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_sobol_points_with_seeds.argtypes = [C.POINTER(A.SobolEntry), C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.oracle_sobol_seeds.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_pmath.argtypes = [C.c_int, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return L


def render(built, spp=None, seed=0x5EED, sample_range=None, tile=None, threads=0, glibc=False):
    spp = spp or built.spp
    o = A.RenderOpts()
    o.seed, o.spp = seed, spp
    o.sampler = getattr(built, "sampler", 0)
    o.sample_begin, o.sample_end = sample_range if sample_range else (0, spp)
    o.tile_x0, o.tile_y0, o.tile_x1, o.tile_y1 = tile if tile else (0, 0, built.width, built.height)
    W, H, Cn = built.width, built.height, built.channels
    block = np.zeros((H, W, Cn, 2), np.float64); light = np.zeros((H, W, Cn), np.float64)
    st = OracleStats()
    rc = (lib_glibc() if glibc else lib()).oracle_render(C.byref(built.desc), C.byref(o), block.ctypes.data_as(C.c_void_p), light.ctypes.data_as(C.c_void_p), threads, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    return block, light, {n: getattr(st, n) for n, _ in st._fields_}
