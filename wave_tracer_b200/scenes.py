"""Procedural restatements of the reference's benchmark scenes (LFS meshes/textures are pointer stubs -- SURVEY.md fact 3).

double_slits(): scenes/diffraction_simple/double_slits.xml + bits/geometry.xml with their default parameters
(all geometry procedural rectangles), `pattern=true` sensor.  The reference file uses plt_bdpt; the same scene is driven
here by plt_path forward + UTD as scenes/diffraction_simple/double_slits_and_reflectors.xml does.
"""
import math
import numpy as np
from .scene import *  # noqa: F401,F403

MM = 1e-3


def double_slits(res=1024, spp=32, direction="forward", max_depth=16, fsd=True, screen=True, lam_mm=.05, with_directional=True,
                 ray_trace_only=False, rr=False, integrator="plt_path", lut=(2048, 1024)):
    L, Lscale, S, E, extent, D, Hh, Z, W, Wslit = -500.0, 1633.0, 50.0, 5.0, 250.0, 12.0, 20.0, -15.0, .65, .35
    lam = lam_mm * MM
    sc = Scene()
    if integrator == "plt_bdpt":     # the reference file's own integrator: <integrator type="plt_bdpt"><integer name="max_depth" value="16"/>
        sc.integrator = PltBdpt(max_depth=max_depth, fsd=fsd, lut=lut)
    else:
        sc.integrator = PltPath(max_depth=max_depth, direction=direction, fsd=fsd, russian_roulette=rr)
    film = Film(res, res // 4, [Discrete(lam)], rfilter_scale=.05)
    sc.sensor = VirtualPlane(lookat((0, 0, (S - .0001) * MM), (0, 0, E * MM), (0, -1, 0)), (extent * MM, extent / 4 * MM), film,
                             alpha=math.radians(.001), samples=spp, ray_trace_only=ray_trace_only)
    # the file's <lookat> has no `up`: the loader takes the tangent of build_orthogonal_frame(dir) = (1,0,0) for dir = +z (transform_loader.cpp:74-76)
    sc.add_emitter(Spot(lookat((0, 0, L * MM), (0, 0, 0), (1, 0, 0)), Discrete(lam, Lscale), cutoff_angle=math.radians(.2), beam_width=math.radians(.1)))
    if with_directional:
        sc.add_emitter(Directional(Blackbody(5750, 1e-6), lookat((-2, 3.5, -1), (0, 0, 0), (1, 0, 0))))
        sc.add_emitter(Directional(Blackbody(6500, 6e-5), lookat((-1, 4, 1), (0, 0, 0), (1, 0, 0))))
    um = 1e-6
    mat_screen = TwoSided(SurfaceSPM(IOR=complex(1, 100), profile=Fractal(.3, gamma=3)))
    mat_floor = TwoSided(Composite([(300e-9, 800e-9, Diffuse(.5)), (1 * um, 1.0, Diffuse(.1))]))
    mat_wall = TwoSided(Diffuse(Binned([(300e-9, 800e-9, .539479), (1 * um, 1.0, .9)])))
    def rect(p, x, y, m):
        sc.add_shape(rectangle(np.array(p) * MM, np.array(x) * MM, np.array(y) * MM), m)
    rect((-100, -Hh, S), (200, 0, 0), (0, 2 * Hh, 0), mat_wall)
    rect((-100, -Hh, L - 100), (200, 0, 0), (0, 0, S - L + 100), mat_floor)
    if screen:
        rect((-D / 2, -Hh, Z), (D / 2 - (W + Wslit) / 2, 0, 0), (0, 2 * Hh, 0), mat_screen)
        rect((-W / 2 + Wslit / 2, -Hh, Z), (W - Wslit, 0, 0), (0, 2 * Hh, 0), mat_screen)
        rect(((W + Wslit) / 2, -Hh, Z), (D / 2 - (W + Wslit) / 2, 0, 0), (0, 2 * Hh, 0), mat_screen)
    return sc


def rgb_response():
    """Three smooth sensitivity curves over 400-700 nm (a stand-in for the reference's RGB response, src/sensor/response/RGB.cpp, whose
    colour-matching tables come from data/sensitivity/XYZ.yml): what matters on the hot path is film_t::splat's per-channel response->f(c, k)
    (film.hpp:254-288) over a polychromatic wavenumber distribution."""
    lam = np.linspace(400e-9, 700e-9, 31)
    g = lambda mu, sg: np.exp(-.5 * ((lam - mu) / sg) ** 2)
    return [Table(lam, 1.0 * g(600e-9, 40e-9) + .35 * g(445e-9, 20e-9)), Table(lam, g(550e-9, 45e-9)), Table(lam, 1.7 * g(450e-9, 25e-9))]


def cornell_like(res=256, spp=16, max_depth=8, lam_nm=550.0, ray_trace_only=False, fsd=False, n_sphere=16, integrator="plt_path", lut=(512, 256), cube_profile=None, rgb=False):
    """A texture-free, procedural cornell-box variant (scenes/cornell-box/box.xml with its PLY shapes dropped):
    5 diffuse walls, a dielectric sphere, a rough-conductor cube, a cube area emitter; perspective sensor; plt_path backward.
    rgb=True: three-channel film over the visible spectrum, 6500 K blackbody emitter, spectrally varying wall reflectances."""
    lam = lam_nm * 1e-9
    sc = Scene()
    sc.integrator = PltBdpt(max_depth=max_depth, fsd=fsd, lut=lut) if integrator == "plt_bdpt" else \
        PltPath(max_depth=max_depth, direction="backward", fsd=fsd, russian_roulette=True)
    film = Film(res, res, rgb_response() if rgb else [Discrete(lam)], rfilter_scale=1.0)
    sc.sensor = Perspective(lookat((0, 1.0, 3.4), (0, 1.0, 0), (0, 1, 0)), math.radians(40), film, ray_trace_only=ray_trace_only, samples=spp)
    white, red, green = TwoSided(Diffuse(.6)), TwoSided(Diffuse(.35)), TwoSided(Diffuse(.45))
    if rgb:
        wl = np.array([400e-9, 480e-9, 520e-9, 580e-9, 620e-9, 700e-9])
        red, green = TwoSided(Diffuse(Table(wl, [.05, .05, .08, .25, .6, .65]))), TwoSided(Diffuse(Table(wl, [.06, .15, .5, .35, .1, .06])))
    sc.add_shape(rectangle((-1, 0, -1), (2, 0, 0), (0, 0, 2)), white)          # floor
    sc.add_shape(rectangle((-1, 2, -1), (0, 0, 2), (2, 0, 0)), white)          # ceiling
    sc.add_shape(rectangle((-1, 0, -1), (0, 2, 0), (2, 0, 0)), white)          # back
    sc.add_shape(rectangle((-1, 0, -1), (0, 0, 2), (0, 2, 0)), red)            # left
    sc.add_shape(rectangle((1, 0, -1), (0, 2, 0), (0, 0, 2)), green)           # right
    sc.add_shape(sphere(.35, (-.4, .35, .2), n_sphere, 2 * n_sphere), Dielectric(1.5))
    sc.add_shape(cube(translate((.45, .3, -.25)) @ rotate((0, 1, 0), .4) @ scale(.3)), SurfaceSPM(IOR=complex(.2, 3.0), profile=cube_profile or Fractal(.2)))
    sc.add_shape(cube(translate((0, 1.98, 0)) @ scale((.25, .01, .25))), Diffuse(.0), emitter=Area(Blackbody(6500, 2e-12) if rgb else Discrete(lam, 1.0), scale=20.0))
    return sc


C0 = 2.99792458e8


def etoile_like(res=720, spp=1024, freq_ghz=10.0, n_buildings=562, seed=7, max_depth=16, fsd=True, ray_trace_only=False, direction="forward"):
    """scenes/sionna_etoile/etoile.xml restated with procedural geometry (its 563 PLY meshes are Git-LFS stubs -- SURVEY.md fact 3, 8d C4).

    Kept from the file: plt_path forward, max_depth 16, russian_roulette off (:24-29); virtual_plane sensor 840 m x 630 m at z = 1 mm, y flipped,
    alpha .001 deg, film res x .75 res, rfilter_scale .1, monochromatic at `wavelength` (:36-62); point emitter at (80.1, 193.8, 21) m with
    phase_space_extent_scale .75 (:194-199); the five ITU materials as twosided(composite(optical diffuse | radio surface_spm, transmission 0))
    (:118-190); the ground plane with mat-itu_concrete (:234-238).  SYNTHETIC: the buildings -- `n_buildings` extruded boxes (12 triangles each)
    on a seeded street plan with a central plaza and twelve radial avenues -- stand in for the 562 building meshes.
    """
    lam = C0 / (freq_ghz * 1e9)
    rng = np.random.default_rng(seed)
    sc = Scene()
    sc.integrator = PltPath(max_depth=max_depth, direction=direction, fsd=fsd, russian_roulette=False)
    film = Film(res, int(res * .75), [Discrete(lam)], rfilter_scale=.1)
    to_world = translate((0, 0, 1e-3)) @ scale((1, -1, 1))
    sc.sensor = VirtualPlane(to_world, (840.0, 630.0), film, alpha=math.radians(.001), samples=spp, ray_trace_only=ray_trace_only)
    em = np.array([80.1, 193.8, 21.0])
    sc.add_emitter(Point(tuple(em), Discrete(lam, 1.0), phase_space_extent_scale=.75))
    rgb = {"marble": (0.701101, 0.644479, 0.485150), "metal": (0.219526, 0.219526, 0.254152), "brick": (0.401968, 0.111874, 0.086764),
           "wood": (0.509804, 0.167376, 0.059954), "concrete": (0.39479, 0.39479, 0.39480)}
    mats = {m: TwoSided(Composite([(300e-9, 800e-9, Diffuse(float(np.mean(c)))), (.1e-3, 1.0, SurfaceSPM(IOR=ITU(m), transmission_scale=0.0))])) for m, c in rgb.items()}
    sc.add_shape(rectangle((-550, -450, 0), (1100, 0, 0), (0, 900, 0)), mats["concrete"])      # mesh-Plane
    # street plan: 30 m cells; plaza of radius 70 m; twelve avenues 22 m wide
    cells = [(x, y) for x in np.arange(-465, 466, 30.0) for y in np.arange(-345, 346, 30.0)]
    keep = []
    for (x, y) in cells:
        r = math.hypot(x, y)
        if r < 70: continue
        ang = math.atan2(y, x) % (math.pi / 6)
        if min(ang, math.pi / 6 - ang) * r < 11: continue
        if math.hypot(x - em[0], y - em[1]) < 28: continue
        keep.append((x, y))
    order = rng.permutation(len(keep))[:n_buildings]
    names = list(rgb)
    for i in sorted(order):
        x, y = keep[i]
        w, d, h = rng.uniform(16, 26), rng.uniform(16, 26), rng.uniform(12, 38)
        M = translate((x + rng.uniform(-2, 2), y + rng.uniform(-2, 2), 0)) @ rotate((0, 0, 1), math.atan2(y, x) + rng.uniform(-.15, .15))
        sc.add_shape(box((-w / 2, -d / 2, 0), (w / 2, d / 2, h), to_world=M), mats[names[int(rng.integers(0, 5))]])
    return sc
