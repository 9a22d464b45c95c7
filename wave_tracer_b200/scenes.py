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


def cornell_like(res=256, spp=16, max_depth=8, lam_nm=550.0, ray_trace_only=False, fsd=False, n_sphere=16, integrator="plt_path", lut=(512, 256), cube_profile=None, rgb=False):
    """A texture-free, procedural cornell-box variant (scenes/cornell-box/box.xml with its PLY shapes dropped):
    5 diffuse walls, a dielectric sphere, a rough-conductor cube, a cube area emitter; perspective sensor; plt_path backward.
    rgb=True: the CIE-RGB / D55 response of box.xml (three channels over 390-830 nm), 6500 K blackbody emitter, spectrally varying wall reflectances."""
    lam = lam_nm * 1e-9
    sc = Scene()
    sc.integrator = PltBdpt(max_depth=max_depth, fsd=fsd, lut=lut) if integrator == "plt_bdpt" else \
        PltPath(max_depth=max_depth, direction="backward", fsd=fsd, russian_roulette=True)
    film = Film(res, res, rgb_response("CIE", "D55") if rgb else [Discrete(lam)], rfilter_scale=1.0)
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


def cornell_box(res=1440, spp=1024, dragon_tris=184320, bunny_tris=81920, lut=(2048, 1024), ray_trace_only=False, defines=None):
    """BASELINE.json configs[0] / [2]: scenes/cornell-box/box.xml as restated in wave_tracer_b200/data/scenes/cornell_box.xml --
    plt_bdpt max_depth 16 with RR, MIS and Fraunhofer FSD; perspective camera with the CIE-RGB / D55 response; CFL spot emitters + 7000 K area
    source over the visible spectrum; Au / Al / SF5 / SF11 materials from the refractive-index tables; prism, flattened sphere, lens, pipe, cube.
    SYNTHETIC (the files are Git-LFS stubs): the dragon (Stanford dragon_vrip_res2: 202,520 triangles, about 0.2 units long before the scene's 4.8x
    scale) is three interpenetrating seeded blobs (body, neck, tail: 81,920 + 81,920 + 20,480 = 184,320 triangles) and the bunny (bun_zipper:
    69,451 triangles, ~0.15 units) one blob of 81,920; the star-of-David screen is an extruded six-pointed star; textures/tiles2.png is the
    constant 0.5 (reflectance 0.35 x 0.5).  About 281k triangles in all (dragon_tris / bunny_tris scale the blobs down for small test renders)."""
    import os
    from . import xml_loader
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "scenes", "cornell_box.xml")
    d = {"res": str(res), "spp": str(spp)}; d.update(defines or {})
    f = dragon_tris / 184320.0
    standins = {"dragon": dict(kind="blob", seed=11, lobes=[(.075, (0, .1, 0), 81920 * f), (.045, (.07, .15, .01), 81920 * f), (.035, (-.09, .07, -.01), 20480 * f)]),
                "bunny_large": dict(kind="blob", radius=.065, centre=(-.02, .1, 0), tris=bunny_tris, seed=23),
                "screen": dict(kind="star", outer=1.0, inner=.55, depth=1.0)}
    sc = xml_loader.load_scene(path, d, lut=lut, missing_meshes="standin", standins=standins, bitmap_standin=.5)
    if ray_trace_only: sc.sensor.rt = True
    return sc


C0 = 2.99792458e8


def extruded_building(footprint, height, floors, to_world=None):
    """A building like the OSM extrusions of the Sionna scenes: a simple polygon footprint (counter-clockwise, local xy) extruded to `height`, its
    walls split into `floors` storeys of two triangles per wall panel, flat roof as a fan.  Face normals."""
    fp = np.asarray(footprint, np.float64); n = len(fp)
    verts, tris = [], []
    def tri(a, b, c):
        t = len(verts); verts.extend([a, b, c]); tris.append((t, t + 1, t + 2))
    zs = np.linspace(0.0, height, floors + 1)
    for i in range(n):
        a, b = fp[i], fp[(i + 1) % n]
        for f in range(floors):
            a0, b0, a1, b1 = (*a, zs[f]), (*b, zs[f]), (*a, zs[f + 1]), (*b, zs[f + 1])
            tri(a0, b0, b1); tri(a0, b1, a1)
    c = fp.mean(axis=0)
    for i in range(n):
        tri((*c, height), (*fp[i], height), (*fp[(i + 1) % n], height))
    return Mesh(np.array(verts, np.float32), np.array(tris, np.uint32), to_world=to_world)


def etoile_like(res=720, spp=1024, freq_ghz=10.0, n_buildings=562, seed=7, max_depth=16, fsd=True, ray_trace_only=False, direction="forward", detail=0):
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
        mat = mats[names[int(rng.integers(0, 5))]]
        if detail <= 0:
            sc.add_shape(box((-w / 2, -d / 2, 0), (w / 2, d / 2, h), to_world=M), mat)
        else:   # detail = storeys per building: ~2 * (10..16) * detail + 13 triangles each (562 buildings at detail 6: ~1e5 triangles)
            nv = int(rng.integers(10, 17))
            ang = np.sort(rng.uniform(0, TWO_PI, nv))
            ang = np.linspace(0, TWO_PI, nv, endpoint=False) + rng.uniform(-.12, .12, nv)
            rad = rng.uniform(.75, 1.0, nv)
            fp = np.stack([.5 * w * rad * np.cos(ang), .5 * d * rad * np.sin(ang)], 1)
            sc.add_shape(extruded_building(fp, h, detail, to_world=M), mat)
    return sc


def _grid_mesh(P, to_world=None, closed_u=False):
    """triangulates a (nu, nv, 3) grid of points; closed_u: the u direction wraps around"""
    nu, nv = P.shape[:2]
    idx = np.arange(nu * nv).reshape(nu, nv)
    iu = np.arange(nu if closed_u else nu - 1); ju = (iu + 1) % nu
    a, b, c, d = idx[iu][:, :-1], idx[ju][:, :-1], idx[ju][:, 1:], idx[iu][:, 1:]
    tris = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)], 0)
    return Mesh(P.reshape(-1, 3).astype(np.float32), tris.astype(np.uint32), to_world=to_world)


def sponza_like(res=1920, spp=4096, max_depth=64, seed=3, detail=1.0):
    """BASELINE.json configs[4]: scenes/sponza/sponza_day.xml -- plt_path backward, max_depth 64 (:8-11); perspective camera, fov 45,
    at (-10, 1, 0) m looking at (0, 2.5, 0), film res x 3/4 res, CIE-RGB / D50 response (:13-30); a 6000 K directional sun (:33-40), an LED bulb
    as a 15 cm area-emitting sphere (:42-55), a 15 m sky panel 60 m up radiating a 7200 K blackbody (:57-71).
    SYNTHETIC: bits/sponza.xml's OBJ mesh (262k triangles, 56 textures: Git-LFS stubs) is replaced by a procedural two-storey atrium of about the
    same triangle count -- floor, walls, two rows of fluted columns carrying arches, an upper gallery, and hanging drapes with folds -- with
    spectrally varying diffuse materials in place of the textures and a few rough-conductor / dielectric details."""
    rng = np.random.default_rng(seed)
    sc = Scene()
    sc.integrator = PltPath(max_depth=max_depth, direction="backward", fsd=True, russian_roulette=True)
    film = Film(res, res * 3 // 4, rgb_response("CIE", "D50"), rfilter_scale=1.0)
    sc.sensor = Perspective(lookat((-10, 1, 0), (0, 2.5, 0), (0, 1, 0)), math.radians(45), film, samples=spp)
    sc.add_emitter(Directional(Blackbody(6000, 2e-3 * 1e-12), lookat((0, 0, 0), (0, -3, -1), (0, 1, 0))))
    wl = np.array([400e-9, 460e-9, 520e-9, 580e-9, 640e-9, 700e-9])
    def paint(r, g, b): return TwoSided(Diffuse(Table(wl, [b * .9, b, g, .5 * (g + r), r, r * .95])))
    stone, floor_m, brick, red, green, blue = paint(.62, .58, .5), paint(.45, .42, .38), paint(.55, .35, .28), paint(.6, .08, .06), paint(.1, .45, .12), paint(.08, .12, .5)
    L, Wd, Hh = 14.0, 6.0, 8.0     # half-length (x), half-width (z), height (y)
    sc.add_shape(rectangle((-L, 0, -Wd), (0, 0, 2 * Wd), (2 * L, 0, 0), tessellation=max(1, int(8 * detail))), floor_m)      # floor (normal +y)
    sc.add_shape(rectangle((-L, 0, -Wd), (2 * L, 0, 0), (0, Hh, 0)), brick)          # wall z = -Wd (normal +z)
    sc.add_shape(rectangle((-L, 0, Wd), (0, Hh, 0), (2 * L, 0, 0)), brick)           # wall z = +Wd (normal -z)
    sc.add_shape(rectangle((L, 0, -Wd), (0, 0, 2 * Wd), (0, Hh, 0)), stone)          # end wall x = +L
    sc.add_shape(rectangle((-L, 0, -Wd), (0, Hh, 0), (0, 0, 2 * Wd)), stone)         # end wall x = -L (behind the camera)
    nu, nv = int(48 * detail), int(24 * detail)
    for storey, (y0, hcol, rcol) in enumerate(((0.0, 3.2, .28), (4.0, 2.6, .2))):
        for side in (-1, 1):
            for ix in range(10):
                x = -L + 1.6 + ix * (2 * L - 3.2) / 9
                z = side * (Wd - 1.7)
                u = np.linspace(0, TWO_PI, nu, endpoint=False)[:, None]; v = np.linspace(0, 1, nv)[None, :]
                flute = 1 + .04 * np.cos(12 * u)                                       # fluted shaft with a slight entasis
                r = rcol * flute * (1 - .12 * v ** 2)
                P = np.stack([x + r * np.cos(u), y0 + hcol * v + 0 * u, z + r * np.sin(u)], -1)
                sc.add_shape(_grid_mesh(P[::-1], closed_u=True), stone)
                if ix < 9:      # arch to the next column: half a torus in the xy-plane
                    xc = x + .5 * (2 * L - 3.2) / 9; R = .5 * (2 * L - 3.2) / 9; r0 = .17
                    th = np.linspace(0, math.pi, int(40 * detail))[:, None]; ph = np.linspace(0, TWO_PI, int(14 * detail), endpoint=False)[None, :]
                    Pa = np.stack([xc - (R + r0 * np.cos(ph)) * np.cos(th), y0 + hcol + (R * .55 + r0 * np.cos(ph)) * np.sin(th), z + r0 * np.sin(ph) + 0 * th], -1)
                    sc.add_shape(_grid_mesh(np.transpose(Pa, (1, 0, 2)), closed_u=True), stone)
        if storey == 0:     # gallery floors over the aisles
            for side in (-1, 1):
                z0 = side * Wd if side < 0 else Wd - 2.4
                sc.add_shape(box((-L, 3.75, min(side * Wd, side * (Wd - 2.4))), (L, 4.0, max(side * Wd, side * (Wd - 2.4)))), stone)
    nd = int(110 * detail)
    for j, (mat, x0) in enumerate(((red, -7.0), (green, -2.5), (blue, 2.0), (red, 6.5), (green, 10.0), (blue, -11.0))):    # drapes hung across the nave
        u = np.linspace(0, 1, nd)[:, None]; v = np.linspace(0, 1, nd)[None, :]
        fold = .18 * np.sin(9 * TWO_PI * u + rng.uniform(0, TWO_PI)) * (1 - .5 * v) + .25 * np.sin(math.pi * u) * v
        P = np.stack([x0 + fold, 7.4 - 3.2 * v + .3 * np.sin(math.pi * u) * (1 - v) + 0 * u, (2 * u - 1) * (Wd - 2.6) + 0 * v], -1)
        sc.add_shape(_grid_mesh(P), mat)
    sc.add_shape(icosphere(.15, (5.75, 1.25, 5.0), 32), Diffuse(.5), emitter=Area(Emission("2723_LED_Greatwall-Ledlight_A19", 2.0)))
    sc.add_shape(square(15.0, to_world=lookat((0, 60, 0), (0, 0, 0), (1, 0, 0))), Diffuse(0.0), emitter=Area(Blackbody(7200, 8e-4 * 1e-12)))
    sc.add_shape(icosphere(.45, (2.0, .45, -1.2), 64), Dielectric(Material("BK7")))                                              # a glass ball and
    sc.add_shape(cube_len(1.0, translate((-2.0, .5, 1.5)) @ rotate((0, 1, 0), .6)), SurfaceSPM(IOR=Material("Cu"), profile=Fractal(.15)))   # a copper block on the floor
    return sc
