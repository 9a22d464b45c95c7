"""Host-side scene model mirroring wave_tracer's plugin interface for the hot path, flattened to wtgpu_scene_desc.

Names and parameters follow the reference's scene elements (XML `type=` strings):
  integrators  plt_path                      (src/integrator/plt_path.cpp:64-94)
  bsdfs        diffuse, dielectric, surface_spm, twosided, composite, scale (src/bsdf/bsdf_loader.cpp:42-57)
  profiles     dirac, fractal, gaussian              (src/interaction/surface_profile/*)
  emitters     point, spot, directional, area (src/emitter/emitter_loader.cpp)
  sensors      perspective, virtual_plane    (src/sensor/sensor_loader.cpp)
  shapes       rectangle, cube, sphere, mesh (src/scene/shape.cpp)
Geometry preparation (triangles, BVH, edges) is done by the native host library (wthost_ads_build);
spectra are baked to constants / uniform tables over the sensor's wavenumber range.
Lengths: metres.  Wavelengths: metres.  Wavenumbers: 1/mm.
"""
import ctypes as C
import math
import numpy as np

from . import _abi as A

TWO_PI = 2.0 * math.pi


def wavelen_to_wavenum(lam_m):
    """k [1/mm] = 2 pi / lambda (include/wt/math/quantity/math.hpp:25-27)."""
    return TWO_PI / (lam_m * 1e3)


# ------------------------------------------------------------------------------------------------ transforms
def lookat(origin, target, up=(0, 1, 0)):
    """transform_t::lookat (include/wt/math/transform/transform.hpp:198-213); returns row-major 4x4 (float64)."""
    o = np.asarray(origin, np.float64)
    d = np.asarray(target, np.float64) - o
    d /= np.linalg.norm(d)
    l = np.cross(np.asarray(up, np.float64), d)
    l /= np.linalg.norm(l)
    u = np.cross(d, l)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = l, u, d, o
    return M


def translate(v):
    M = np.eye(4); M[:3, 3] = v; return M


def scale(v):
    M = np.eye(4); M[0, 0], M[1, 1], M[2, 2] = (v, v, v) if np.isscalar(v) else v; return M


def rotate(axis, angle_rad):
    a = np.asarray(axis, np.float64); a /= np.linalg.norm(a)
    c, s = math.cos(angle_rad), math.sin(angle_rad)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    M = np.eye(4); M[:3, :3] = c * np.eye(3) + s * K + (1 - c) * np.outer(a, a); return M


# ------------------------------------------------------------------------------------------------ spectra
def bake_binned_spectrum(values, kmin, kmax):
    """The tables of binned_piecewise_linear_distribution_t's constructor (binned_piecewise_linear_distribution.hpp:36-63) in ITS arithmetic:
    f32 grid step, f32 running trapezoid sum in index order, one f32 reciprocal of the total, f32 products.  Returns (ys, dcdf, dx, norm, sum),
    bit-identical with the reference's members (DESIGN.md 7's pin table, row `binned_piecewise_linear_distribution.hpp`)."""
    f = np.float32
    ys = np.asarray(values, f); n = len(ys)
    dx = f(f(kmax) - f(kmin)) / f(n - 1)
    terms = (dx * (ys[1:] + ys[:-1])).astype(f) / f(2)
    dcdf = np.concatenate([np.zeros(1, f), np.add.accumulate(terms, dtype=f)]).astype(f)       # accumulate is strictly sequential
    tot = dcdf[-1]
    norm = f(1) / tot if tot > 0 else f(0)
    return ys, (dcdf * norm).astype(f), dx, norm, tot


def bake_discrete_cdf(densities):
    """discrete_distribution_t's constructor (discrete_distribution.hpp:45-66) in its arithmetic: f32 running sum of max(0, density) in index
    order, one f32 reciprocal of the total, f32 products; a distribution without mass gets dcdf.back() = 1.  n + 1 entries."""
    f = np.float32
    d = np.maximum(np.asarray(densities, f), f(0))
    cdf = np.concatenate([np.zeros(1, f), np.add.accumulate(d, dtype=f)]).astype(f)
    if cdf[-1] > 0: return (cdf * (f(1) / cdf[-1])).astype(f)
    cdf[-1] = 1
    return cdf


class Spectrum:
    """spectrum_t / spectrum_real_t (include/wt/spectrum/spectrum.hpp:37-93): value(k) complex, f(k) real."""
    def value(self, k):  # k: ndarray of 1/mm
        raise NotImplementedError


class Const(Spectrum):
    def __init__(self, v): self.v = complex(v)
    def value(self, k): return np.full(np.shape(k), self.v, np.complex128)


class Discrete(Spectrum):
    """spectrum/discrete.hpp: lines (wavelength -> value); f(k) is non-zero only exactly on a line."""
    def __init__(self, wavelength_m, value=1.0):
        self.lines = [(np.float32(wavelen_to_wavenum(wavelength_m)), float(value))]
    def value(self, k):
        out = np.zeros(np.shape(k), np.complex128)
        for kl, v in self.lines:
            out = np.where(np.asarray(k, np.float32) == kl, v, out)
        return out


class Blackbody(Spectrum):
    """spectrum/blackbody.hpp: Planck spectral radiance per unit wavenumber (arbitrary but fixed normalisation), times scale."""
    def __init__(self, T, scale=1.0): self.T, self.scale = float(T), float(scale)
    def value(self, k):
        k = np.asarray(k, np.float64)
        lam = TWO_PI / np.maximum(k, 1e-30) * 1e-3          # m
        h, c, kb = 6.62607015e-34, 2.99792458e8, 1.380649e-23
        nu = c / lam
        B = 2 * h * nu ** 3 / c ** 2 / np.expm1(np.minimum(h * nu / (kb * self.T), 700.0))
        return (self.scale * B * (c / TWO_PI) * 1e3).astype(np.complex128)   # per (1/mm)


class Table(Spectrum):
    """piecewise-linear spectrum over wavelengths (spectrum/piecewise_linear.hpp); zero outside."""
    def __init__(self, wavelengths_m, values):
        k = wavelen_to_wavenum(np.asarray(wavelengths_m, np.float64))
        o = np.argsort(k); self.k, self.v = k[o], np.asarray(values, np.complex128)[o]
    def value(self, k):
        k = np.asarray(k, np.float64)
        re = np.interp(k, self.k, self.v.real, left=0, right=0); im = np.interp(k, self.k, self.v.imag, left=0, right=0)
        return re + 1j * im


class Binned(Spectrum):
    """composite spectrum (spectrum/composite.hpp): wavelength bins -> spectrum; zero outside all bins."""
    def __init__(self, bins): self.bins = [(wavelen_to_wavenum(hi), wavelen_to_wavenum(lo), s) for lo, hi, s in bins]
    def value(self, k):
        k = np.asarray(k, np.float64); out = np.zeros(k.shape, np.complex128)
        for kmin, kmax, s in self.bins:
            m = (k >= kmin) & (k < kmax)
            if m.any(): out[m] = _as_spectrum(s).value(k[m])       # a bin's spectrum is only asked about its own range
        return out


class ITU(Spectrum):
    """Complex IOR of an ITU-R P.2040-2 (table 3) material: sqrt(eps_r - i sigma / (eps0 omega)), eps_r = a f^b, sigma = c f^d, f in GHz
    (src/spectrum/util/spectrum_from_ITU.cpp:31-55, table :78-170).  Zero outside the material's frequency range."""
    TABLE = {"vacuum": [(1, 0, 0, 0, 0, float("inf"))], "concrete": [(5.24, 0, 0.0462, 0.7822, 1, 100)], "brick": [(3.91, 0, 0.0238, 0.16, 1, 40)],
             "plasterboard": [(2.73, 0, 0.0085, 0.9395, 1, 100)], "wood": [(1.99, 0, 0.0047, 1.0718, 0.001, 100)],
             "glass": [(6.31, 0, 0.0036, 1.3394, 0.1, 100), (5.79, 0, 0.0004, 1.658, 220, 450)], "chipboard": [(2.58, 0, 0.0217, 0.78, 1, 100)],
             "plywood": [(2.71, 0, 0.33, 0, 1, 40)], "marble": [(7.074, 0, 0.0055, 0.9262, 1, 60)], "metal": [(1, 0, 1e7, 0, 1, 100)]}
    def __init__(self, material): self.params = self.TABLE[material]
    def value(self, k):
        k = np.asarray(k, np.float64); c0, eps0 = 2.99792458e8, 8.8541878128e-12
        omega = k * 1e3 * c0; f_ghz = omega / TWO_PI * 1e-9
        out = np.zeros(k.shape, np.complex128)
        for a, b, c, d, f0, f1 in self.params:
            eps = a * (f_ghz ** b if b else 1.0); sig = c * (f_ghz ** d if d else 1.0)
            v = np.sqrt(eps - 1j * sig / (eps0 * np.maximum(omega, 1e-300)))
            out = np.where((f_ghz >= f0 * (1 - 1e-6)) & (f_ghz <= f1 * (1 + 1e-6)), v, out)      # endpoints inclusive up to f32 rounding of k
        return out


_SPECTRA = None


def spectra_db():
    """wave_tracer_b200/data/spectra.npz: the tabulated physical data the reference scenes name (refractive indices of data/ior, the emission
    spectrum box.xml uses, the CIE colour-matching functions), extracted by tools/extract_reference_spectra.py."""
    global _SPECTRA
    if _SPECTRA is None:
        import os
        _SPECTRA = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spectra.npz"))
    return _SPECTRA


def Material(name):
    """<spectrum name="IOR" material="Au"/>: complex refractive index n + i k, piecewise linear in wavenumber (src/spectrum/util/spectrum_from_db.cpp:60-200)."""
    db = spectra_db()
    if f"ior/{name}/n" not in db: raise ValueError(f"material {name!r} is not in the spectra fixture (tools/extract_reference_spectra.py)")
    return Table(db[f"ior/{name}/lam_um"].astype(np.float64) * 1e-6, db[f"ior/{name}/n"].astype(np.float64) + 1j * db[f"ior/{name}/k"].astype(np.float64))


class Scaled(Spectrum):
    def __init__(self, s, scale): self.s, self.scale = _as_spectrum(s), float(scale)
    def value(self, k): return self.s.value(k) * self.scale


def Emission(name, scale=1.0):
    """<spectrum emitter="2534_CFL_Tensor_Twister">: a tabulated lamp spectrum of data/emission (LSPDD), times scale."""
    db = spectra_db()
    if f"emission/{name}/value" not in db: raise ValueError(f"emission spectrum {name!r} is not in the spectra fixture")
    return Scaled(Table(db[f"emission/{name}/lam_nm"].astype(np.float64) * 1e-9, db[f"emission/{name}/value"].astype(np.float64)), scale)


def rgb_response(colourspace="CIE", white_point="D55"):
    """<response type="RGB">: f(channel, k) = max(0, (M xyz(k))[channel]) with the CIE colour-matching functions and the XYZ -> RGB matrix of the
    colourspace, Bradford-adapted to the white point (src/sensor/response/RGB.cpp:30-42; spectrum/colourspace/RGB/RGB.hpp:39-130;
    whitepoint.hpp:28-61).  Returns three Table spectra."""
    M = {"CIE": ("E", [[2.3706743, -0.9000405, -0.4706338], [-0.5138850, 1.4253036, 0.0885814], [0.0052982, -0.0146949, 1.0093968]]),
         "sRGB": ("D65", [[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]])}[colourspace]
    W = {"D50": (0.96422, 1.0, 0.82521), "D55": (0.95682, 1.0, 0.92149), "D65": (0.95047, 1.0, 1.08883), "D75": (0.94972, 1.0, 1.22638), "E": (1.0, 1.0, 1.0)}
    X2R = np.array(M[1])
    if M[0] != white_point:      # Bradford chromatic adaptation
        MA = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]])
        iMA = np.array([[0.9869929, -0.1470543, 0.1599627], [0.4323053, 0.5183603, 0.0492912], [-0.0085287, 0.0400428, 0.9684867]])
        rs, rd = MA @ np.array(W[M[0]]), MA @ np.array(W[white_point])
        X2R = X2R @ (iMA @ np.diag(rd / rs) @ MA)
    db = spectra_db()
    lam = db["XYZ/lam_nm"].astype(np.float64) * 1e-9
    rgb = np.maximum(db["XYZ/xyz"].astype(np.float64) @ X2R.T, 0.0)
    return [Table(lam, rgb[:, c]) for c in range(3)]


def _as_spectrum(s):
    return s if isinstance(s, Spectrum) else Const(s)


# ------------------------------------------------------------------------------------------------ bsdfs
class Bsdf: pass


class Diffuse(Bsdf):
    def __init__(self, reflectance): self.reflectance = _as_spectrum(reflectance)


class Dielectric(Bsdf):
    def __init__(self, IOR, extIOR=1.0, reflection_scale=None, transmission_scale=None):
        self.IOR, self.extIOR, self.rs, self.ts = _as_spectrum(IOR), _as_spectrum(extIOR), reflection_scale, transmission_scale


class Dirac: pass


class Fractal:
    def __init__(self, roughness, gamma=3.0): self.roughness, self.gamma = _as_spectrum(roughness), float(gamma)


class Gaussian:
    """surface_profile type="gaussian" (interaction/surface_profile/gaussian.hpp): exactly one of `roughness` (perceptual) or `sigma` (1/mm)."""
    def __init__(self, roughness=None, sigma=None):
        if (roughness is None) == (sigma is None): raise ValueError("gaussian surface profile: either 'roughness' or 'sigma' must be provided")   # gaussian.cpp:55-56
        self.roughness = None if roughness is None else _as_spectrum(roughness)
        self.sigma = None if sigma is None else _as_spectrum(sigma)


class SurfaceSPM(Bsdf):
    def __init__(self, IOR, extIOR=1.0, profile=None, reflection_scale=None, transmission_scale=None):
        self.IOR, self.extIOR, self.profile = _as_spectrum(IOR), _as_spectrum(extIOR), profile or Dirac()
        self.rs, self.ts = reflection_scale, transmission_scale


class TwoSided(Bsdf):
    def __init__(self, nested): self.nested = nested


class Scale(Bsdf):
    def __init__(self, scale, nested): self.scale, self.nested = _as_spectrum(scale), nested


class Composite(Bsdf):
    """bins: list of (wavelength_min_m, wavelength_max_m, bsdf) (src/bsdf/composite.cpp:42-95)."""
    def __init__(self, bins): self.bins = bins


# ------------------------------------------------------------------------------------------------ emitters / sensors / shapes
class Point:
    def __init__(self, position, radiant_intensity, extent=None, phase_space_extent_scale=1.0):
        self.position, self.spectrum, self.extent, self.pse = position, _as_spectrum(radiant_intensity), extent, phase_space_extent_scale


class Spot:
    """src/emitter/spot.cpp:82-133: cutoff_angle default 20 deg, beam_width default .75 cutoff."""
    def __init__(self, to_world, radiant_intensity, cutoff_angle=math.radians(20), beam_width=None, extent=None, phase_space_extent_scale=1.0):
        self.to_world, self.spectrum = np.asarray(to_world, np.float64), _as_spectrum(radiant_intensity)
        self.cutoff = float(cutoff_angle); self.falloff = float(beam_width) if beam_width is not None else .75 * self.cutoff
        self.extent, self.pse = extent, phase_space_extent_scale


class Directional:
    """src/emitter/directional.cpp:86-132: dir = to_world * (0,0,-1) is the direction TO the emitter."""
    def __init__(self, irradiance, to_world=None, solid_angle=6.794e-5, phase_space_extent_scale=1.0):
        d = np.array([0, 0, -1.0])
        if to_world is not None:
            d = np.asarray(to_world, np.float64)[:3, :3] @ d; d /= np.linalg.norm(d)
        self.dir, self.spectrum, self.solid_angle, self.pse = d, _as_spectrum(irradiance), solid_angle, phase_space_extent_scale


class Area:
    def __init__(self, radiance, scale=1.0, phase_space_extent_scale=1.0):
        self.spectrum, self.scale, self.pse = _as_spectrum(radiance), float(scale), phase_space_extent_scale


class Film:
    """film_t (include/wt/sensor/film/film.hpp:116-135, 355-420); response: list of per-channel spectra."""
    def __init__(self, width, height, response, rfilter_scale=1.0):
        self.width, self.height, self.response, self.rfilter_scale = int(width), int(height), [_as_spectrum(r) for r in response], float(rfilter_scale)


class VirtualPlane:
    def __init__(self, to_world, extent, film, alpha=None, ray_trace_only=False, samples=1):
        self.to_world, self.extent, self.film, self.alpha, self.rt, self.samples = np.asarray(to_world, np.float64), extent, film, alpha, ray_trace_only, samples


class Perspective:
    def __init__(self, to_world, fov, film, ray_trace_only=False, samples=1, sourcing_tan_alpha=None, phase_space_extent_scale=1.0):
        self.to_world, self.fov, self.film, self.rt, self.samples = np.asarray(to_world, np.float64), float(fov), film, ray_trace_only, samples
        self.sta, self.pse = sourcing_tan_alpha, phase_space_extent_scale


class PltPath:
    def __init__(self, max_depth=1024, direction="backward", fsd=True, russian_roulette=True):
        self.max_depth, self.direction, self.fsd, self.rr = int(max_depth), direction, bool(fsd), bool(russian_roulette)


class PltBdpt:
    """plt_bdpt options (src/integrator/plt_bdpt.cpp:161-197); lut=(n, m) sizes of the regenerated Fraunhofer sampling tables."""
    def __init__(self, max_depth=1024, fsd=True, russian_roulette=True, mis=True, sensor_direct_sampling=True, emitter_direct_sampling=True, lut=(2048, 1024)):
        self.max_depth, self.fsd, self.rr, self.mis = int(max_depth), bool(fsd), bool(russian_roulette), bool(mis)
        self.sensor_direct, self.emitter_direct, self.lut = bool(sensor_direct_sampling), bool(emitter_direct_sampling), lut


class Mesh:
    def __init__(self, positions, indices, normals=None, uvs=None, to_world=None):
        self.positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        self.normals = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        self.uvs = None if uvs is None else np.ascontiguousarray(uvs, np.float32).reshape(-1, 2)
        self.to_world = np.eye(4) if to_world is None else np.asarray(to_world, np.float64)


def rectangle(p, x, y, to_world=None, tessellation=1):
    """src/mesh/rectangle.cpp:26-79: 4 vertices + 2 triangles per cell, uv = cell corners."""
    p, x, y = (np.asarray(v, np.float32) for v in (p, x, y))
    verts, uvs, tris = [], [], []
    rt = np.float32(1.0) / np.float32(tessellation)
    for ix in range(tessellation):
        for iy in range(tessellation):
            t = len(verts)
            u0, v0 = np.float32(ix) * rt, np.float32(iy) * rt
            u1 = np.float32(1) if ix + 1 == tessellation else np.float32(ix + 1) * rt
            v1 = np.float32(1) if iy + 1 == tessellation else np.float32(iy + 1) * rt
            for (u, v) in ((u0, v0), (u1, v0), (u1, v1), (u0, v1)):
                verts.append(p + u * x + v * y); uvs.append((u, v))
            tris += [(t, t + 1, t + 2), (t + 2, t + 3, t)]
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), uvs=np.array(uvs, np.float32), to_world=to_world)


def cube(to_world=None):
    """unit cube [-1,1]^3 with per-face vertices (outward normals)."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    verts, uvs, tris = [], [], []
    for n, a, b in faces:
        n, a, b = np.array(n, np.float32), np.array(a, np.float32), np.array(b, np.float32)
        t = len(verts)
        for (u, v) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            verts.append(n + u * a + v * b); uvs.append(((u + 1) / 2, (v + 1) / 2))
        tris += [(t, t + 1, t + 2), (t + 2, t + 3, t)]
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), uvs=np.array(uvs, np.float32), to_world=to_world)


def box(lo, hi, to_world=None):
    """axis-aligned box [lo, hi] as a transformed cube (src/mesh/cube.cpp): 12 triangles, face normals."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    M = translate((lo + hi) / 2) @ scale(tuple((hi - lo) / 2))
    return cube(M if to_world is None else to_world @ M)


def sphere(radius=1.0, centre=(0, 0, 0), n_lat=16, n_lon=32, to_world=None):
    verts, normals, uvs, tris = [], [], [], []
    for i in range(n_lat + 1):
        th = math.pi * i / n_lat
        for j in range(n_lon + 1):
            ph = TWO_PI * j / n_lon
            n = np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
            verts.append(np.asarray(centre) + radius * n); normals.append(n); uvs.append((j / n_lon, i / n_lat))
    for i in range(n_lat):
        for j in range(n_lon):
            a, b = i * (n_lon + 1) + j, (i + 1) * (n_lon + 1) + j
            if i > 0: tris.append((a, b, a + 1))
            if i < n_lat - 1: tris.append((a + 1, b, b + 1))
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), normals=np.array(normals, np.float32), uvs=np.array(uvs, np.float32), to_world=to_world)


def square(length, to_world=None, tessellation=1):
    """<shape type="rectangle"> with `length`: a square in the local xy-plane centred at the origin (src/mesh/rectangle.cpp:74-89)."""
    return rectangle((-length / 2, -length / 2, 0), (length, 0, 0), (0, length, 0), to_world=to_world, tessellation=tessellation)


def cube_len(length=2.0, to_world=None):
    """<shape type="cube"> with `length` (default 2 m): side `length`, centred at the origin (src/mesh/cube.cpp)."""
    M = scale(length / 2)
    return cube(M if to_world is None else np.asarray(to_world, np.float64) @ M)


def prism(length=1.0, height=1.0, angle=math.pi / 2, to_world=None):
    """<shape type="prism">: triangular prism along z, base on y = 0 of width height*tan(angle/2), apex at y = height (src/mesh/prism.cpp): 8 triangles."""
    hx, hz = .5 * height * math.tan(angle / 2), .5 * length
    A0, B0, C0 = (-hx, 0, -hz), (hx, 0, -hz), (0, height, -hz)
    A1, B1, C1 = (-hx, 0, hz), (hx, 0, hz), (0, height, hz)
    verts, tris = [], []
    def face(ps, uv):
        t = len(verts); verts.extend(ps)
        if len(ps) == 3: tris.append((t, t + 1, t + 2))
        else: tris.extend([(t, t + 1, t + 2), (t + 2, t + 3, t)])
    face([A0, C0, B0], None); face([A1, B1, C1], None)          # end caps (outward: -z, +z)
    face([A1, C1, C0, A0], None); face([B0, C0, C1, B1], None)  # slanted sides
    face([A1, A0, B0, B1], None)                                # base
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), to_world=to_world)


def icosphere(radius=1.0, centre=(0, 0, 0), tessellation=32, to_world=None, displace=None):
    """<shape type="sphere">: an icosahedron subdivided round(log2(tessellation/3)) times, per-face vertices with radial normals
    (src/mesh/sphere.cpp:20-94, icosahedron.cpp): 20 * 4^r triangles (1280 at the default tessellation 32).
    displace(unit_dirs) -> radii: optional radial displacement (procedural stand-ins for scanned meshes); normals are then recomputed per face."""
    a, b = 1.0, 2.0 / (1.0 + math.sqrt(5.0))
    V = np.array([(0, b, -a), (b, a, 0), (-b, a, 0), (0, b, a), (0, -b, a), (-a, 0, b), (0, -b, -a), (a, 0, -b), (a, 0, b), (-a, 0, -b), (b, -a, 0), (-b, -a, 0)], np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array([(2, 1, 0), (1, 2, 3), (5, 4, 3), (4, 8, 3), (7, 6, 0), (6, 9, 0), (11, 10, 4), (10, 11, 6), (9, 5, 2), (5, 9, 11), (8, 7, 1), (7, 8, 10),
                  (2, 5, 3), (8, 1, 3), (9, 2, 0), (1, 7, 0), (11, 9, 6), (7, 10, 6), (5, 11, 4), (10, 8, 4)], np.int64)
    tri = V[F]                                                   # (20, 3, 3)
    rec = int(max(0.0, math.log2(tessellation / 3.0)) + .5)
    for _ in range(rec):
        p0, p1, p2 = tri[:, 0], tri[:, 1], tri[:, 2]
        p01, p02, p12 = (p0 + p1) / 2, (p0 + p2) / 2, (p1 + p2) / 2
        tri = np.concatenate([np.stack([p0, p01, p02], 1), np.stack([p01, p1, p12], 1), np.stack([p01, p12, p02], 1), np.stack([p02, p12, p2], 1)], 0)
    n = tri.reshape(-1, 3); n = n / np.linalg.norm(n, axis=1, keepdims=True)
    r = radius if displace is None else radius * np.asarray(displace(n), np.float64)[:, None]
    pos = n * r + np.asarray(centre, np.float64)
    idx = np.arange(len(pos), dtype=np.uint32).reshape(-1, 3)
    normals = n
    if displace is not None:
        fn = np.cross(pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]]); fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-30)
        normals = np.repeat(fn, 3, axis=0)
    return Mesh(pos.astype(np.float32), idx, normals=normals.astype(np.float32), to_world=to_world)


def blob(radius, centre, n_tris, seed=1, roughness=.35, to_world=None):
    """SYNTHETIC stand-in for a scanned mesh (the reference's dragon / bunny PLYs are Git-LFS stubs): an icosphere of ~n_tris triangles whose
    radius is modulated by a few seeded low-frequency lobes -- a closed, smooth-shaded, non-convex surface with the triangle budget of the original."""
    rec = max(0, int(round(math.log(max(n_tris, 20) / 20.0, 4))))
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=(9, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    amp, freq, ph = rng.uniform(.3, 1.0, 9) * roughness / 3, rng.integers(2, 6, 9), rng.uniform(0, TWO_PI, 9)
    def displace(n):
        d = np.ones(len(n))
        for a, f, p, w in zip(amp, freq, ph, dirs): d += a * np.sin(f * np.arccos(np.clip(n @ w, -1, 1)) + p)
        return np.maximum(d, .3)
    return icosphere(radius, centre, tessellation=3 * 2 ** rec, to_world=to_world, displace=displace)


def cylinder(p0, p1, radius, tessellation=32, to_world=None):
    """<shape type="cylinder">: an open tube from p0 to p1, 2 * tessellation triangles, radial normals (src/mesh/cylinder.cpp)."""
    p0, p1 = np.asarray(p0, np.float64), np.asarray(p1, np.float64)
    v = p1 - p0; L = np.linalg.norm(v); n = v / L
    if abs(n[0]) > abs(n[1]): x = 1 / math.sqrt(n[0] ** 2 + n[2] ** 2); b = np.array([x * n[2], 0, -x * n[0]])
    else: x = 1 / math.sqrt(n[1] ** 2 + n[2] ** 2); b = np.array([0, x * n[2], -x * n[1]])
    t = np.cross(b, n)
    verts, normals, uvs, tris = [], [], [], []
    for i in range(tessellation):
        phi = TWO_PI * i / tessellation; c, s_ = math.cos(phi), math.sin(phi)
        d = c * t + s_ * b
        verts += [p0 + radius * d, p0 + radius * d + L * n]; normals += [d, d]; uvs += [(i / tessellation, 0), (i / tessellation, 1)]
        i0, i1, i2 = 2 * i, 2 * i + 1, (2 * i + 2) % (2 * tessellation); i3 = i2 + 1
        tris += [(i0, i2, i1), (i1, i2, i3)]
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), normals=np.array(normals, np.float32), uvs=np.array(uvs, np.float32), to_world=to_world)


def lens(radius, centre, R1, R2, thickness=0.0, tessellation=50, to_world=None):
    """<shape type="lens">: two spherical caps of curvature R1 / R2 (in units of 1/radius; 0: flat; sign: convex / concave) facing -x / +x, rim radius
    `radius`, joined by a cylindrical edge when the edge thickness is positive (src/mesh/lens.cpp).  Rings are spaced as (i/n)^0.8 like the
    reference's; ~4 tessellation^2 triangles."""
    Rl = radius / R1 if R1 != 0 else math.inf
    Rr = radius / R2 if R2 != 0 else math.inf
    x1 = math.copysign(math.sqrt(Rl * Rl - radius * radius), Rl) if math.isfinite(Rl) else 0.0
    x2 = -math.copysign(math.sqrt(Rr * Rr - radius * radius), Rr) if math.isfinite(Rr) else 0.0
    ET = x1 - x2 - (Rl if math.isfinite(Rl) else 0.0) - (Rr if math.isfinite(Rr) else 0.0) + thickness
    if thickness == 0 and R1 <= 0 and R2 <= 0: ET += radius / 1000
    verts, normals, tris = [], [], []
    def cap(R, xc, sign, shift):
        start = len(verts)
        nt = tessellation if math.isfinite(R) else 1
        apex = np.array([xc - sign * (R if math.isfinite(R) else 0.0) + shift, 0, 0]) if sign < 0 else np.array([xc + (R if math.isfinite(R) else 0.0) + shift, 0, 0])
        verts.append(apex); normals.append(np.array([sign, 0, 0], np.float64))
        for i in range(nt):
            h = radius * min(1.0, ((i + 1) / nt) ** .8)
            for j in range(tessellation):
                phi = TWO_PI * j / tessellation
                cp = np.array([0, math.cos(phi) * h, math.sin(phi) * h])
                if math.isfinite(R):
                    c0 = np.array([xc, 0, 0]); n = cp - c0; n /= np.linalg.norm(n)
                    if R < 0: n = -n
                    p = c0 + n * R + np.array([shift, 0, 0])
                else:
                    n = np.array([sign, 0, 0], np.float64); p = cp + np.array([shift, 0, 0])
                verts.append(p); normals.append(n)
        for i in range(nt):
            for j in range(tessellation):
                jp = j - 1 if j > 0 else tessellation - 1
                cur, prv = start + 1 + i * tessellation + j, start + 1 + i * tessellation + jp
                if i == 0: t = (start, cur, prv)
                else:
                    a0, a1 = start + 1 + (i - 1) * tessellation + jp, start + 1 + (i - 1) * tessellation + j
                    tris.append((a0, a1, prv) if sign < 0 else (a1, a0, prv)); t = (prv, a1, cur) if sign < 0 else (a1, prv, cur)
                tris.append(t if sign < 0 or i > 0 else (start, prv, cur))
    cap(Rl, x1, -1, 0.0)
    cap(Rr, x2, +1, ET)
    if ET > 0:
        e0 = len(verts)
        for j in range(tessellation):
            phi = TWO_PI * j / tessellation; n = np.array([0, math.cos(phi), math.sin(phi)])
            verts += [n * radius, n * radius + np.array([ET, 0, 0])]; normals += [n, n]
        for j in range(tessellation):
            p0 = 2 * j - 2 if j > 0 else 2 * tessellation - 2
            tris += [(e0 + p0 + 1, e0 + p0, e0 + 2 * j), (e0 + 2 * j + 1, e0 + p0 + 1, e0 + 2 * j)]
    V = np.array(verts, np.float64) + np.asarray(centre, np.float64)
    return Mesh(V.astype(np.float32), np.array(tris, np.uint32), normals=np.array(normals, np.float32), to_world=to_world)


def star_prism(outer=1.0, inner=.5, depth=.1, points=6, to_world=None):
    """SYNTHETIC stand-in for box.xml's star_big.ply ("star of david", loaded with face normals): a flat `points`-pointed star in the local yz-plane,
    extruded along x by `depth`: 4 * 2 * points triangles, face normals (sharp edges everywhere -- it is the scene's diffracting screen)."""
    ang = np.arange(2 * points) * math.pi / points
    rad = np.where(np.arange(2 * points) % 2 == 0, outer, inner)
    ring = np.stack([np.zeros_like(ang), rad * np.cos(ang), rad * np.sin(ang)], 1)
    verts, tris = [], []
    def tri(a, b, c):
        t = len(verts); verts.extend([a, b, c]); tris.append((t, t + 1, t + 2))
    for side, x in ((-1, -depth / 2), (1, depth / 2)):
        c = np.array([x, 0, 0])
        for i in range(2 * points):
            a, b = ring[i] + c, ring[(i + 1) % (2 * points)] + c
            tri(c, a, b) if side > 0 else tri(c, b, a)
    for i in range(2 * points):
        a, b = ring[i], ring[(i + 1) % (2 * points)]
        a0, b0, a1, b1 = a + [-depth / 2, 0, 0], b + [-depth / 2, 0, 0], a + [depth / 2, 0, 0], b + [depth / 2, 0, 0]
        tri(a0, b1, a1); tri(a0, b0, b1)
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), to_world=to_world)


# ------------------------------------------------------------------------------------------------ scene + flattening
class BuiltScene:
    """Owns the ctypes wtgpu_scene_desc and everything it points to."""
    def __init__(self):
        self.desc = A.SceneDesc()
        self.keep = []
        self.ads = None
        self.sampler = 0        # WTGPU_SAMPLER_*
    def __del__(self):
        try:
            if self.ads is not None: A.host_lib().wthost_ads_destroy(self.ads)
        except Exception:
            pass
    @property
    def width(self): return self.desc.sensor.width
    @property
    def height(self): return self.desc.sensor.height
    @property
    def channels(self): return self.desc.sensor.channels


def _arr(ctype, values):
    a = (ctype * max(1, len(values)))(*values)
    return a


class Sobolld:
    """<sampler type="sobolld"> (src/sampler/sobolld.cpp:84-99): the scene sampler; `table` = parsed initIrreducibleGF3.dat entries
    (wave_tracer_b200/sobol.py), default: the real file if `path` holds one, else the stand-in table."""
    def __init__(self, table=None, path=None):
        from . import sobol
        self.table = table if table is not None else sobol.table_for(path)


class Scene:
    def __init__(self):
        self.sampler = None     # None: sampler::uniform_t (src/scene/loader/loader.cpp:299-300); or Sobolld()
        self.integrator = PltPath()
        self.sensor = None
        self.shapes = []        # (mesh, bsdf, area_emitter or None)
        self.emitters = []      # non-area emitters

    def add_shape(self, mesh, bsdf, emitter=None):
        self.shapes.append((mesh, bsdf, emitter)); return len(self.shapes) - 1

    def add_emitter(self, e):
        self.emitters.append(e); return e

    # -- sensor wavenumber range (sensitivity_spectrum().wavenumber_range())
    def _krange(self):
        ks = []
        for r in self.sensor.film.response:
            if isinstance(r, Discrete): ks += [k for k, _ in r.lines]
            elif isinstance(r, Table): ks += [r.k[0], r.k[-1]]
            else: raise ValueError("sensor response must be Discrete or Table")
        return float(min(ks)), float(max(ks))

    def build(self, table_size=256):
        L = A.host_lib()
        out = BuiltScene()
        d = out.desc
        d.api_version = 1
        kmin, kmax = self._krange()
        mono = kmin == kmax

        spectra, spectrum_data = [], []
        def bake(s):
            s = _as_spectrum(s)
            sp = A.Spectrum()
            if mono or isinstance(s, Const):
                v = complex(s.value(np.array([np.float32(kmin)]))[0]) if not isinstance(s, Const) else s.v
                sp.type, sp.re, sp.im = A.SPECTRUM_CONSTANT, v.real, v.imag
            else:
                ks = np.linspace(kmin, kmax, table_size)
                v = s.value(ks)
                sp.type, sp.k0, sp.inv_dk, sp.n, sp.offset = A.SPECTRUM_TABLE, kmin, (table_size - 1) / (kmax - kmin), table_size, len(spectrum_data) // 2
                for c in v: spectrum_data.extend((float(c.real), float(c.imag)))
            spectra.append(sp); return len(spectra) - 1

        bsdfs, bins = [], []
        flat_ids = {}          # a bsdf object shared between shapes (the XML's <ref id=...>) is flattened once
        def flat_bsdf(b):
            if id(b) in flat_ids: return flat_ids[id(b)]
            flat_ids[id(b)] = len(bsdfs)
            n = A.Bsdf(); n.child = -1; n.spec[:] = [-1] * 4; n.prof_spec[:] = [-1] * 2
            idx = len(bsdfs); bsdfs.append(n)
            if isinstance(b, Diffuse):
                n.type = A.BSDF_DIFFUSE; n.spec[0] = bake(b.reflectance)
            elif isinstance(b, (Dielectric, SurfaceSPM)):
                n.type = A.BSDF_DIELECTRIC if isinstance(b, Dielectric) else A.BSDF_SURFACE_SPM
                n.spec[0], n.spec[1] = bake(b.extIOR), bake(b.IOR)
                if b.rs is not None: n.spec[2] = bake(b.rs)
                if b.ts is not None: n.spec[3] = bake(b.ts)
                if isinstance(b, SurfaceSPM):
                    if isinstance(b.profile, Fractal):
                        n.profile_type, n.gamma = A.PROFILE_FRACTAL_ROUGHNESS, b.profile.gamma; n.prof_spec[0] = bake(b.profile.roughness)
                    elif isinstance(b.profile, Gaussian):
                        if b.profile.roughness is not None: n.profile_type = A.PROFILE_GAUSSIAN; n.prof_spec[0] = bake(b.profile.roughness)
                        else: n.profile_type = A.PROFILE_GAUSSIAN_SIGMA; n.prof_spec[0] = bake(b.profile.sigma)
                    else:
                        n.profile_type = A.PROFILE_DIRAC
            elif isinstance(b, TwoSided):
                n.type = A.BSDF_TWO_SIDED; n.child = flat_bsdf(b.nested)
            elif isinstance(b, Scale):
                n.type = A.BSDF_SCALE; n.spec[0] = bake(b.scale); n.child = flat_bsdf(b.nested)
            elif isinstance(b, Composite):
                n.type = A.BSDF_COMPOSITE
                # bins that cannot be selected at any wavenumber the sensor queries are dropped (composite.hpp:64-70 picks a bin by k; k stays
                # inside the sensor's range): their spectra need not be defined there (e.g. rgb reflectances in a microwave scene)
                live = [(lo, hi, c) for lo, hi, c in b.bins if wavelen_to_wavenum(hi) <= kmax and kmin < wavelen_to_wavenum(lo)]
                children = [(wavelen_to_wavenum(hi), wavelen_to_wavenum(lo), flat_bsdf(c)) for lo, hi, c in live]
                n.bin_first, n.n_bins = len(bins), len(children)
                for kmn, kmx, c in children:
                    bb = A.BsdfBin(); bb.kmin, bb.kmax, bb.child = kmn, kmx, c; bins.append(bb)
            else:
                raise ValueError(f"unsupported bsdf {type(b)}")
            return idx

        # ---- emitters (area emitters are appended after the others, in shape order)
        emitters = []
        for e in self.emitters:
            emitters.append((e, -1))
        meshes = (A.MeshDesc * len(self.shapes))()
        for si, (mesh, bsdf, em) in enumerate(self.shapes):
            m = meshes[si]
            m.n_verts = len(mesh.positions); m.positions = mesh.positions.ctypes.data_as(A.P(A.c_f))
            m.normals = mesh.normals.ctypes.data_as(A.P(A.c_f)) if mesh.normals is not None else None
            m.uvs = mesh.uvs.ctypes.data_as(A.P(A.c_f)) if mesh.uvs is not None else None
            m.n_tris = len(mesh.indices); m.indices = mesh.indices.ctypes.data_as(A.P(A.c_u32))
            m.to_world[:] = list(np.asarray(mesh.to_world, np.float64).reshape(-1))
            m.bsdf = flat_bsdf(bsdf)
            m.emitter = -1
            if em is not None:
                m.emitter = len(emitters); emitters.append((em, si))
            out.keep.append(mesh)
        ads = C.c_void_p()
        A.check_host(L.wthost_ads_build(len(self.shapes), meshes, C.byref(ads)), "wthost_ads_build")
        out.ads = ads
        A.check_host(L.wthost_ads_fill(ads, C.byref(d)), "wthost_ads_fill")
        out.keep.append(meshes)

        wmin, wmax = np.array(d.world_min[:]), np.array(d.world_max[:])
        # ---- sensor
        s = d.sensor
        film = self.sensor.film
        s.width, s.height, s.channels = film.width, film.height, len(film.response)
        stddev = np.float32(.25) * np.float32(film.rfilter_scale)     # beam_source_spatial_stddev * rfilter_scale (film.hpp:417)
        s.rfilter_stddev = float(stddev)
        s.rf_radius = int(np.uint32(np.float32(math.ceil(float(stddev * np.float32(3)))) + np.float32(.5)))
        s.ray_trace_only = 1 if self.sensor.rt else 0
        s.response[:] = [bake(r) for r in film.response] + [-1] * (4 - len(film.response))
        M = self.sensor.to_world
        R = M[:3, :3]
        if isinstance(self.sensor, VirtualPlane):
            s.type = A.SENSOR_VIRTUAL_PLANE
            t, b, n = (R[:, i] / np.linalg.norm(R[:, i]) for i in range(3))
            ext = np.array(self.sensor.extent, np.float64) * np.array([np.linalg.norm(R[:, 0]), np.linalg.norm(R[:, 1])])
            centre = M[:3, 3]
            origin = centre - ext[0] / 2 * t - ext[1] / 2 * b
            s.frame_t[:], s.frame_b[:], s.frame_n[:], s.origin[:], s.extent[:] = list(t), list(b), list(n), list(origin), list(ext)
            s.requested_tan_alpha = math.tan(self.sensor.alpha) if self.sensor.alpha is not None else -1.0
        else:
            s.type = A.SENSOR_PERSPECTIVE
            s.pos[:] = list(M[:3, 3]); s.rot[:] = list(R.reshape(-1)); s.inv_rot[:] = list(np.linalg.inv(R).reshape(-1))
            W, H = film.width, film.height
            h = 1.0 / math.tan(self.sensor.fov / 2); w = h / (W / H); znear = 0.01
            Pm = np.zeros((4, 4)); Pm[0, 0], Pm[1, 1], Pm[3, 2], Pm[2, 3] = w, h, 1.0, znear
            V = scale((.5 * (W - 1), .5 * (H - 1), 1)) @ translate((1, 1, 0)) @ scale((-1, -1, 1))
            S2C = np.linalg.inv(V @ Pm)
            s.s2c[:] = list(S2C.reshape(-1)); s.c2s[:] = list((V @ Pm).reshape(-1))
            if self.sensor.sta is not None:
                s.sourcing_tan_alpha = self.sensor.sta
            else:   # pixel_len / image_plane_z (perspective.hpp:147-151)
                def pos(fx, fy):
                    p = S2C @ np.array([fx, fy, 1, 1.0]); return p[:3] / p[3]
                s.sourcing_tan_alpha = float(np.linalg.norm(pos(W, 0) - pos(0, 0)) / W / znear)
            s.pse_scale = self.sensor.pse
            s.requested_tan_alpha = -1.0

        # ---- emitters + spectral sampling tables (scene_build_sensor_sampling_data.cpp:40-150, simplified product spectra)
        em_structs, powers, kdists, kdist_data = [], [], [], []
        sens = lambda k: sum(_as_spectrum(r).value(k).real for r in film.response)
        ktab = np.array([kmin], np.float64) if mono else np.linspace(kmin, kmax, 1024)
        shapes_area = [d.shapes[i].surface_area for i in range(d.n_shapes)]
        centre = (wmin + wmax) / 2; pr = (wmax - wmin) / 2
        for e, si in emitters:
            E = A.Emitter(); E.shape = -1; E.extent = -1.0
            E.spectrum = bake(e.spectrum); E.pse_scale = e.pse; E.scale = 1.0
            em_f = e.spectrum.value(ktab).real
            if isinstance(e, Point):
                E.type = A.EMITTER_POINT; E.pos[:] = list(e.position); geom = 4 * math.pi
                if e.extent: E.extent = e.extent
            elif isinstance(e, Spot):
                E.type = A.EMITTER_SPOT; Mw = e.to_world; E.pos[:] = list(Mw[:3, 3])
                E.rot[:] = list(Mw[:3, :3].reshape(-1)); E.inv_rot[:] = list(np.linalg.inv(Mw[:3, :3]).reshape(-1))
                E.cutoff, E.falloff = e.cutoff, e.falloff
                if e.extent: E.extent = e.extent
                geom = TWO_PI * (1 - .5 * (math.cos(e.cutoff) + math.cos(e.falloff)))
            elif isinstance(e, Directional):
                E.type = A.EMITTER_DIRECTIONAL; E.dir[:] = list(e.dir)
                E.tan_alpha = math.tan(math.acos(1 - e.solid_angle / TWO_PI))
                # set_world_aabb (directional.hpp:56-82)
                n = e.dir
                if abs(n[0]) > abs(n[1]):
                    x = 1 / math.sqrt(n[0] ** 2 + n[2] ** 2); b = np.array([x * n[2], 0, -x * n[0]])
                else:
                    x = 1 / math.sqrt(n[1] ** 2 + n[2] ** 2); b = np.array([0, x * n[2], -x * n[1]])
                t = np.cross(b, n)
                r2, far = 0.0, 0.0
                for sy in (-1, 1):
                    for sz in (-1, 1):
                        c = np.array([-pr[0], sy * pr[1], sz * pr[2]])
                        loc = np.array([c @ t, c @ b, c @ n]); r2 = max(r2, loc[0] ** 2 + loc[1] ** 2); far = max(far, abs(loc[2]))
                E.world_centre[:] = list(centre); E.world_radius = math.sqrt(r2); E.far_dist = 1.01 * far
                geom = math.pi * r2
            elif isinstance(e, Area):
                E.type = A.EMITTER_AREA; E.shape = si; E.scale = e.scale
                geom = e.scale * shapes_area[si] * math.pi
            else:
                raise ValueError(f"unsupported emitter {type(e)}")
            prod = np.maximum(em_f * sens(ktab), 0)
            kd = A.KDist(); kd.first = len(kdist_data)
            if mono:
                kd.type, kd.n, kd.norm = A.KDIST_DISCRETE, 1, (1.0 / prod[0] if prod[0] > 0 else 0.0)
                kdist_data += [float(np.float32(kmin)), float(prod[0]), 0.0, 1.0]
                power = geom * prod[0]
            else:
                n = len(ktab)
                ys, dcdf, dk, norm, tot = bake_binned_spectrum(prod, kmin, kmax)
                kd.type, kd.n, kd.k0, kd.dk, kd.norm = A.KDIST_BINNED, n, kmin, float(dk), float(norm)
                kdist_data += [float(v) for v in ys] + [float(v) for v in dcdf]
                power = geom * float(tot)
            em_structs.append(E); powers.append(power); kdists.append(kd)
        if not em_structs:
            raise ValueError("(scene) no emitters defined")
        # discrete_distribution_t power cdf, float32 accumulate (discrete_distribution.hpp:45-66)
        tot = sum(powers)
        cdf = list(bake_discrete_cdf([p / tot if tot > 0 else 0 for p in powers]))

        # ---- tables into desc
        def put(name_n, name_p, ctype, values, keepname):
            arr = _arr(ctype, values); out.keep.append(arr)
            if name_n: setattr(d, name_n, len(values))
            setattr(d, name_p, arr)
        put("n_spectra", "spectra", A.Spectrum, spectra, "spectra")
        put("n_spectrum_data", "spectrum_data", A.c_f, spectrum_data, "sd")
        put("n_bsdfs", "bsdfs", A.Bsdf, bsdfs, "bsdfs")
        put("n_bsdf_bins", "bsdf_bins", A.BsdfBin, bins, "bins")
        put("n_emitters", "emitters", A.Emitter, em_structs, "em")
        put(None, "emitter_cdf", A.c_f, [float(c) for c in cdf], "cdf")
        put(None, "emitter_kdist", A.KDist, kdists, "kd")
        put("n_kdist_data", "kdist_data", A.c_f, kdist_data, "kdd")
        # shapes' bsdf/emitter ids were recorded by wthost_ads_build from the mesh descs

        it = d.integrator
        it.max_depth, it.russian_roulette, it.fsd = self.integrator.max_depth, int(self.integrator.rr), int(self.integrator.fsd)
        if isinstance(self.integrator, PltBdpt):
            it.type = A.INTEGRATOR_PLT_BDPT
            it.mis, it.sensor_direct, it.emitter_direct = int(self.integrator.mis), int(self.integrator.sensor_direct), int(self.integrator.emitter_direct)
            if self.integrator.fsd and not self.sensor.rt:      # plt_bdpt.cpp:189-194: the LUTs are only loaded when not ray tracing
                from . import fsd_lut
                n, m = self.integrator.lut
                t1, t2, c1, c2 = (np.ascontiguousarray(x, np.float32) for x in fsd_lut.build(n, m))
                out.keep += [t1, t2, c1, c2]
                d.fsd_lut_n, d.fsd_lut_m = n, m
                d.fsd_icdf_theta1, d.fsd_icdf_theta2 = t1.ctypes.data_as(A.P(A.c_f)), t2.ctypes.data_as(A.P(A.c_f))
                d.fsd_icdf1, d.fsd_icdf2 = c1.ctypes.data_as(A.P(A.c_f)), c2.ctypes.data_as(A.P(A.c_f))
        else:
            it.type = A.INTEGRATOR_PLT_PATH
            it.direction = A.DIRECTION_FORWARD if self.integrator.direction == "forward" else A.DIRECTION_BACKWARD
        out.spp = self.sensor.samples
        out.sampler = A.SAMPLER_UNIFORM
        if isinstance(self.sampler, Sobolld):
            from . import sobol
            arr = sobol.to_abi(self.sampler.table); out.keep.append(arr)
            d.sobol_table = arr
            out.sampler = A.SAMPLER_SOBOLLD
        return out
