"""ctypes mirror of include/wtgpu.h and include/wthost.h (the C-ABI drop-in boundary).

Layout must match the C structs field for field; tests/test_abi.py cross-checks every sizeof against the
library (wtgpu_debug_sizeof).  The product library is loaded from wave_tracer_b200/libwt_b200.so and the
import FAILS LOUDLY when it is missing: there is no CPU fallback on the product path.
"""
import ctypes as C
import os

c_f, c_u32, c_i32, c_u64, c_dbl = C.c_float, C.c_uint32, C.c_int32, C.c_uint64, C.c_double
INVALID_IDX = 0xFFFFFFFF
SAMPLER_UNIFORM, SAMPLER_SOBOLLD = 0, 1


class Node(C.Structure):
    _fields_ = [("minx", c_f * 8), ("miny", c_f * 8), ("minz", c_f * 8), ("maxx", c_f * 8), ("maxy", c_f * 8),
                ("maxz", c_f * 8), ("child", c_i32 * 8), ("tris_start", c_u32), ("tris_count", c_u32), ("pad_", c_u32 * 6)]


class Leaf(C.Structure):
    _fields_ = [("tris_ptr", c_u32), ("count", c_u32)]


class Tri(C.Structure):
    _fields_ = [(n, c_f) for n in ("ax", "ay", "az", "nx", "bx", "by", "bz", "ny", "cx", "cy", "cz", "nz")]


class TriMeta(C.Structure):
    _fields_ = [("shape_idx", c_u32), ("shape_tri_idx", c_u32), ("edge_ab", c_u32), ("edge_bc", c_u32), ("edge_ca", c_u32), ("pad_", c_u32 * 3)]


class TriShading(C.Structure):
    _fields_ = [("n0", c_f * 3), ("n1", c_f * 3), ("n2", c_f * 3), ("uv0", c_f * 2), ("uv1", c_f * 2), ("uv2", c_f * 2), ("dpdu", c_f * 3), ("has_uv", c_u32)]


class Edge(C.Structure):
    _fields_ = [("a", c_f * 3), ("b", c_f * 3), ("e", c_f * 3), ("n1", c_f * 3), ("t1", c_f * 3), ("n2", c_f * 3), ("t2", c_f * 3),
                ("alpha", c_f), ("tri1", c_u32), ("tri2", c_u32)]


class Shape(C.Structure):
    _fields_ = [("bsdf", c_i32), ("emitter", c_i32), ("surface_area", c_f), ("tri_first", c_u32), ("n_tris", c_u32), ("cdf_first", c_u32), ("pad_", c_u32 * 2)]


SPECTRUM_CONSTANT, SPECTRUM_TABLE = 0, 1


class Spectrum(C.Structure):
    _fields_ = [("type", c_u32), ("re", c_f), ("im", c_f), ("k0", c_f), ("inv_dk", c_f), ("n", c_u32), ("offset", c_u32), ("pad_", c_u32)]


BSDF_DIFFUSE, BSDF_DIELECTRIC, BSDF_SURFACE_SPM, BSDF_TWO_SIDED, BSDF_COMPOSITE, BSDF_SCALE, BSDF_MASK = range(7)
PROFILE_DIRAC, PROFILE_GAUSSIAN, PROFILE_FRACTAL_ROUGHNESS, PROFILE_FRACTAL_T, PROFILE_GAUSSIAN_SIGMA = range(5)


class Bsdf(C.Structure):
    _fields_ = [("type", c_u32), ("child", c_i32), ("spec", c_i32 * 4), ("profile_type", c_u32), ("prof_spec", c_i32 * 2), ("gamma", c_f),
                ("n_bins", c_u32), ("bin_first", c_u32), ("pad_", c_u32 * 3)]


class BsdfBin(C.Structure):
    _fields_ = [("kmin", c_f), ("kmax", c_f), ("child", c_i32), ("pad_", c_u32)]


EMITTER_POINT, EMITTER_SPOT, EMITTER_DIRECTIONAL, EMITTER_AREA = range(4)


class Emitter(C.Structure):
    _fields_ = [("type", c_u32), ("spectrum", c_i32), ("scale", c_f), ("pse_scale", c_f), ("pos", c_f * 3), ("rot", c_f * 9), ("inv_rot", c_f * 9),
                ("cutoff", c_f), ("falloff", c_f), ("extent", c_f), ("shape", c_i32), ("dir", c_f * 3), ("tan_alpha", c_f),
                ("world_centre", c_f * 3), ("world_radius", c_f), ("far_dist", c_f), ("pad_", c_u32 * 2)]


KDIST_DISCRETE, KDIST_BINNED = 0, 1


class KDist(C.Structure):
    _fields_ = [("type", c_u32), ("n", c_u32), ("first", c_u32), ("k0", c_f), ("dk", c_f), ("norm", c_f), ("pad_", c_u32 * 2)]


SENSOR_PERSPECTIVE, SENSOR_VIRTUAL_PLANE = 0, 1


class Sensor(C.Structure):
    _fields_ = [("type", c_u32), ("width", c_u32), ("height", c_u32), ("channels", c_u32), ("rfilter_stddev", c_f), ("rf_radius", c_u32),
                ("ray_trace_only", c_u32), ("response", c_i32 * 4),
                ("pos", c_f * 3), ("rot", c_f * 9), ("inv_rot", c_f * 9), ("s2c", c_f * 16), ("c2s", c_f * 16), ("sourcing_tan_alpha", c_f), ("pse_scale", c_f),
                ("frame_t", c_f * 3), ("frame_b", c_f * 3), ("frame_n", c_f * 3), ("origin", c_f * 3), ("extent", c_f * 2), ("requested_tan_alpha", c_f),
                ("pad_", c_u32 * 2)]


INTEGRATOR_PLT_PATH, INTEGRATOR_PLT_BDPT = 0, 1
DIRECTION_BACKWARD, DIRECTION_FORWARD = 0, 1


class Integrator(C.Structure):
    _fields_ = [("type", c_u32), ("direction", c_u32), ("max_depth", c_u32), ("russian_roulette", c_u32), ("fsd", c_u32),
                ("mis", c_u32), ("sensor_direct", c_u32), ("emitter_direct", c_u32)]


P = C.POINTER


class SobolEntry(C.Structure):
    _fields_ = [("d", c_i32), ("sj", c_i32), ("aj", c_i32), ("mk", c_i32 * 32), ("pad_", c_i32)]


class SceneDesc(C.Structure):
    _fields_ = [("api_version", c_u32),
                ("n_nodes", c_u32), ("nodes", P(Node)), ("n_leaves", c_u32), ("leaves", P(Leaf)), ("root_ptr", c_i32),
                ("n_tris", c_u32), ("tris", P(Tri)), ("tri_meta", P(TriMeta)), ("tri_shading", P(TriShading)),
                ("n_edges", c_u32), ("edges", P(Edge)), ("world_min", c_f * 3), ("world_max", c_f * 3),
                ("n_shapes", c_u32), ("shapes", P(Shape)), ("n_shape_tris", c_u32), ("shape_tri_tuid", P(c_u32)),
                ("n_shape_cdf", c_u32), ("shape_tri_cdf", P(c_f)),
                ("n_spectra", c_u32), ("spectra", P(Spectrum)), ("n_spectrum_data", c_u32), ("spectrum_data", P(c_f)),
                ("n_bsdfs", c_u32), ("bsdfs", P(Bsdf)), ("n_bsdf_bins", c_u32), ("bsdf_bins", P(BsdfBin)),
                ("n_emitters", c_u32), ("emitters", P(Emitter)), ("emitter_cdf", P(c_f)), ("emitter_kdist", P(KDist)),
                ("n_kdist_data", c_u32), ("kdist_data", P(c_f)),
                ("sensor", Sensor), ("integrator", Integrator),
                ("fsd_lut_n", c_u32), ("fsd_lut_m", c_u32), ("fsd_icdf_theta1", P(c_f)), ("fsd_icdf_theta2", P(c_f)), ("fsd_icdf1", P(c_f)), ("fsd_icdf2", P(c_f)),
                ("sobol_table", P(SobolEntry))]


class RenderOpts(C.Structure):
    _fields_ = [("seed", c_u64), ("spp", c_u32), ("sample_begin", c_u32), ("sample_end", c_u32),
                ("tile_x0", c_u32), ("tile_y0", c_u32), ("tile_x1", c_u32), ("tile_y1", c_u32),
                ("device", c_i32), ("film_on_device", c_u32), ("pool_size", c_u32), ("sampler", c_u32), ("flags", c_u32), ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [(n, c_u64) for n in ("samples", "segments", "ray_casts", "cone_casts", "shadow_casts", "nodes_visited", "tris_tested", "edges_fetched",
                                     "surface_interactions", "fsd_interactions", "null_interactions", "splats", "capacity_overflows", "kernel_launches", "iterations",
                                     "traverse_nodes", "traverse_tris", "shaded_paths")] + \
               [(n, c_dbl) for n in ("gpu_ms", "traverse_ms", "shade_ms", "generate_ms", "sort_ms", "connect_ms")] + [("strategies", c_u64 * 5), ("walker_steps", c_u64)] + \
               [(n, c_u32) for n in ("passes", "pool_used", "cap_tris", "cap_edges", "cap_segments", "cap_apertures", "cap_vertices", "subpools")] + [("stack_drops", c_u64)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "strategies" else getattr(self, n)) for n, _ in self._fields_}


class RayQuery(C.Structure):
    _fields_ = [("o", c_f * 3), ("d", c_f * 3), ("tmin", c_f), ("tmax", c_f)]


class RayHit(C.Structure):
    _fields_ = [("tuid", c_u32), ("dist", c_f), ("bary", c_f * 2), ("front_face", c_u32)]


MAX_CONE_TRIS, MAX_CONE_EDGES = 128, 48


class ConeQuery(C.Structure):
    _fields_ = [("o", c_f * 3), ("d", c_f * 3), ("x", c_f * 3), ("x0", c_f), ("tan_alpha", c_f), ("e", c_f), ("tmin", c_f), ("tmax", c_f), ("z_scale", c_f)]


class ConeHit(C.Structure):
    _fields_ = [("dist", c_f), ("front_face", c_u32), ("n_tris", c_u32), ("n_edges", c_u32), ("tris", c_u32 * MAX_CONE_TRIS), ("edges", c_u32 * MAX_CONE_EDGES)]


class MeshDesc(C.Structure):
    _fields_ = [("n_verts", c_u32), ("positions", P(c_f)), ("normals", P(c_f)), ("uvs", P(c_f)), ("n_tris", c_u32), ("indices", P(c_u32)),
                ("to_world", c_dbl * 16), ("bsdf", c_i32), ("emitter", c_i32)]


ABI_STRUCTS = [Node, Leaf, Tri, TriMeta, TriShading, Edge, Shape, Spectrum, Bsdf, BsdfBin, Emitter, KDist, Sensor, Integrator,
               SceneDesc, RenderOpts, Stats, RayQuery, RayHit, ConeQuery, ConeHit, MeshDesc, SobolEntry]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwt_b200.so")
if os.environ.get("WT_B200_LIB"):      # A/B builds of the same sources (tools/gpu_session_*.sh); still the CUDA library, never a fallback
    LIB_PATH = os.path.abspath(os.environ["WT_B200_LIB"])
_lib = None


def lib():
    """Loads the product library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"wave_tracer_b200: native library {LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback on the product path.")
    L = C.CDLL(LIB_PATH)
    L.wtgpu_device_count.restype = C.c_int
    L.wtgpu_last_error.restype = C.c_char_p
    L.wtgpu_scene_create.argtypes = [P(SceneDesc), C.c_int, P(C.c_void_p)]
    L.wtgpu_scene_destroy.argtypes = [C.c_void_p]
    L.wtgpu_scene_destroy.restype = None
    L.wtgpu_trim.argtypes = []; L.wtgpu_trim.restype = None
    L.wtgpu_render.argtypes = [C.c_void_p, P(RenderOpts), C.c_void_p, C.c_void_p, P(Stats)]
    L.wtgpu_get_capacities.argtypes = [C.c_void_p, P(c_u32)]
    L.wtgpu_set_capacities.argtypes = [C.c_void_p, P(c_u32)]
    L.wtgpu_develop.argtypes = [P(Sensor), c_u32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.wtgpu_develop_device.argtypes = [P(Sensor), c_u32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.wtgpu_debug_intersect_rays.argtypes = [C.c_void_p, c_u32, P(RayQuery), P(RayHit)]
    L.wtgpu_debug_shadow_rays.argtypes = [C.c_void_p, c_u32, P(RayQuery), P(c_u32)]
    L.wtgpu_debug_intersect_cones.argtypes = [C.c_void_p, c_u32, P(ConeQuery), P(ConeHit)]
    L.wtgpu_debug_rng.argtypes = [c_u64, c_u32, c_u32, c_u32, P(c_f), C.c_int]
    L.wtgpu_debug_pmath.argtypes = [C.c_int, c_u32, P(c_f), P(c_f), P(c_f), C.c_int]
    L.wtgpu_debug_sobol.argtypes = [C.c_void_p, c_u64, c_u64, c_u32, P(c_u32), P(c_f)]
    L.wtgpu_debug_sizeof.argtypes = [C.c_int]
    L.wtgpu_debug_sizeof.restype = C.c_uint64
    _lib = L
    return L


HOST_LIB_PATH = os.path.join(_HERE, "libwt_host.so")
_host_lib = None


def host_lib():
    """Loads the host-only library (wthost_*: mesh -> triangles -> BVH -> edges, sobolld tables).  It contains no CUDA code and does not load the
    CUDA library: the CPU legs of bench.py (`--impl reference`, cpu_baseline) build their scene tables with it and never map libwt_b200.so."""
    global _host_lib
    if _host_lib is not None:
        return _host_lib
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError(f"wave_tracer_b200: native library {HOST_LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = C.CDLL(HOST_LIB_PATH)
    L.wthost_sobol_tables.argtypes = [P(SobolEntry), P(C.c_uint16), P(C.c_uint16)]
    L.wthost_ads_build.argtypes = [c_u32, P(MeshDesc), P(C.c_void_p)]
    L.wthost_ads_fill.argtypes = [C.c_void_p, P(SceneDesc)]
    L.wthost_ads_destroy.argtypes = [C.c_void_p]
    L.wthost_ads_destroy.restype = None
    L.wthost_ads_sah_cost.argtypes = [C.c_void_p]
    L.wthost_ads_sah_cost.restype = c_dbl
    L.wthost_ads_max_depth.argtypes = [C.c_void_p]
    L.wthost_ads_max_depth.restype = c_u32
    _host_lib = L
    return L


def check_host(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"wthost: {what} failed with code {rc}")


EXPORTED_SYMBOLS = ["wtgpu_device_count", "wtgpu_last_error", "wtgpu_scene_create", "wtgpu_scene_destroy", "wtgpu_trim", "wtgpu_render", "wtgpu_get_capacities", "wtgpu_set_capacities", "wtgpu_develop", "wtgpu_develop_device",
                    "wtgpu_debug_intersect_rays", "wtgpu_debug_shadow_rays", "wtgpu_debug_intersect_cones", "wtgpu_debug_rng", "wtgpu_debug_pmath", "wtgpu_debug_sobol", "wtgpu_debug_sizeof"]
HOST_EXPORTED_SYMBOLS = ["wthost_sobol_tables", "wthost_ads_build", "wthost_ads_fill", "wthost_ads_destroy", "wthost_ads_sah_cost", "wthost_ads_max_depth"]


def check(rc, what=""):
    if rc != 0:
        msg = lib().wtgpu_last_error()
        raise RuntimeError(f"wtgpu: {what} failed with code {rc}: {msg.decode() if msg else ''}")
