#!/usr/bin/env python
"""Bisects a GPU-vs-oracle film difference on the slit_bench scene (tests/test_xml_loader.py::_slit_bench_api) over its ingredients."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _oracle
from wave_tracer_b200 import (Scene, PltPath, Film, VirtualPlane, Spot, Discrete, Diffuse, SurfaceSPM, Gaussian, Fractal, Dirac, TwoSided, Composite, Binned, rectangle, lookat, render)
MM = 1e-3
def mk(profile, with_floor=True, rt=False, fsd=True, gap=.9, lam_mm=.08, res=192, spp=8, ior=complex(1, 80), max_depth=12):
    lam = lam_mm * MM
    sc = Scene()
    sc.integrator = PltPath(max_depth=max_depth, direction="forward", russian_roulette=False, fsd=fsd)
    film = Film(res, res // 3, [Discrete(lam)], rfilter_scale=.1)
    sc.sensor = VirtualPlane(lookat((0, 0, (40 - .001) * MM), (0, 0, 2 * MM), (0, -1, 0)), (200 * MM, 200 / 3 * MM), film, alpha=math.radians(.002), samples=spp, ray_trace_only=rt)
    sc.add_emitter(Spot(lookat((0, 0, -400 * MM), (0, 0, 0), (1, 0, 0)), Discrete(lam, 900.0), cutoff_angle=math.radians(.3), beam_width=math.radians(.15)))
    mat_screen = TwoSided(SurfaceSPM(IOR=ior, profile=profile))
    mat_floor = TwoSided(Composite([(1e-6, 1.0, Diffuse(.15))]))
    mat_wall = TwoSided(Diffuse(Binned([(300e-9, 800e-9, .5), (1e-6, 1.0, .85)])))
    def rect(p, x, y, m): sc.add_shape(rectangle(np.array(p) * MM, np.array(x) * MM, np.array(y) * MM), m)
    rect((-80, -15, 40), (160, 0, 0), (0, 30, 0), mat_wall)
    if with_floor: rect((-80, -15, -450), (160, 0, 0), (0, 0, 490), mat_floor)
    rect((-6, -15, -12), (6 - gap / 2, 0, 0), (0, 30, 0), mat_screen)
    rect((gap / 2, -15, -12), (6 - gap / 2, 0, 0), (0, 30, 0), mat_screen)
    return sc
cases = {
    "gaussian.25": dict(profile=Gaussian(roughness=.25)),
    "gaussian.25 no floor": dict(profile=Gaussian(roughness=.25), with_floor=False),
    "gaussian.25 rt": dict(profile=Gaussian(roughness=.25), rt=True),
    "gaussian.25 nofsd": dict(profile=Gaussian(roughness=.25), fsd=False),
    "gaussian.25 depth2": dict(profile=Gaussian(roughness=.25), max_depth=2),
    "gaussian.05": dict(profile=Gaussian(roughness=.05)),
    "gaussian sigma 50": dict(profile=Gaussian(sigma=50.0)),
    "fractal.25": dict(profile=Fractal(.25)),
    "dirac": dict(profile=Dirac()),
}
for name, kw in cases.items():
    b = mk(**kw).build()
    blk, lgt, st = render(b, spp=8)
    oblk, olgt, ost = _oracle.render(b, spp=8)
    g = lgt.astype(np.float64)
    l2 = np.linalg.norm(g - olgt) / max(np.linalg.norm(olgt), 1e-300); fl = abs(g.sum() - olgt.sum()) / max(abs(olgt.sum()), 1e-300)
    lit = olgt > 0
    worst = np.argsort(-np.abs(g - olgt).ravel())[:3]
    print(f"{name:24s} rel-L2 {l2:.3e} flux {fl:.3e} sum {olgt.sum():.4e} segs {st['segments']} {ost['segments']} surf {st['surface_interactions']} {ost['surface']} fsd {st['fsd_interactions']} {ost['fsd']} worst {[(int(i), float(g.ravel()[i]), float(olgt.ravel()[i])) for i in worst]}")
