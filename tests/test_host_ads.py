"""Host ADS builder (wthost_ads_build): triangles, 8-wide BVH, deterministic edge classification."""
import math
import numpy as np
from wave_tracer_b200 import Scene, PltPath, Film, VirtualPlane, Spot, Discrete, Diffuse, TwoSided, rectangle, cube, lookat
from wave_tracer_b200 import scenes


def _mini(meshes):
    sc = Scene(); sc.integrator = PltPath(max_depth=2, direction="forward")
    lam = 5e-5
    sc.sensor = VirtualPlane(lookat((0, 0, 1), (0, 0, 0), (0, 1, 0)), (1, 1), Film(8, 8, [Discrete(lam)]))
    sc.add_emitter(Spot(lookat((0, 0, -1), (0, 0, 0)), Discrete(lam, 1.0)))
    for m in meshes: sc.add_shape(m, TwoSided(Diffuse(.5)))
    return sc.build()


def test_rectangle_edges():
    b = _mini([rectangle((0, 0, 0), (1, 0, 0), (0, 1, 0))])
    d = b.desc
    assert d.n_tris == 2 and d.n_edges == 4          # the coplanar diagonal (alpha = pi > 160 deg) is discarded
    for i in range(4):
        e = d.edges[i]
        assert e.tri2 == 0xFFFFFFFF and abs(e.alpha) < 1e-6      # open edges are knife edges (n2 = -n1)
        assert abs(np.dot(e.n1[:], e.e[:])) < 1e-6 and abs(np.dot(e.t1[:], e.e[:])) < 1e-6
    metas = [d.tri_meta[i] for i in range(2)]
    assert sum(int(x != 0xFFFFFFFF) for m in metas for x in (m.edge_ab, m.edge_bc, m.edge_ca)) == 4


def test_cube_edges_and_bvh():
    b = _mini([cube()])
    d = b.desc
    assert d.n_tris == 12
    # per-face vertices are duplicated but bit-identical positions are shared: 12 cube edges, each a 90-degree wedge
    alphas = sorted(round(d.edges[i].alpha, 4) for i in range(d.n_edges))
    assert d.n_edges == 12 and all(abs(a - math.pi / 2) < 1e-3 for a in alphas)
    for i in range(d.n_edges):
        assert d.edges[i].tri2 != 0xFFFFFFFF
    assert abs(d.shapes[0].surface_area - 24.0) < 1e-4
    # every triangle is inside its leaf's AABB and node ranges are contiguous
    for ni in range(d.n_nodes):
        n = d.nodes[ni]
        for c in range(8):
            ch = n.child[c]
            if ch < 0:
                lf = d.leaves[-ch - 1]
                for t in range(lf.tris_ptr, lf.tris_ptr + lf.count):
                    tr = d.tris[t]
                    for (x, y, z) in ((tr.ax, tr.ay, tr.az), (tr.bx, tr.by, tr.bz), (tr.cx, tr.cy, tr.cz)):
                        assert n.minx[c] <= x <= n.maxx[c] and n.miny[c] <= y <= n.maxy[c] and n.minz[c] <= z <= n.maxz[c]
                assert n.tris_start <= lf.tris_ptr and lf.tris_ptr + lf.count <= n.tris_start + n.tris_count
    seen = sorted(d.shape_tri_tuid[i] for i in range(d.n_shape_tris))
    assert seen == list(range(12))


def test_builder_is_deterministic():
    a = scenes.cornell_like(res=16, spp=1).build(); b = scenes.cornell_like(res=16, spp=1).build()
    assert a.desc.n_edges == b.desc.n_edges and a.desc.n_nodes == b.desc.n_nodes
    for i in range(a.desc.n_tris):
        ma, mb = a.desc.tri_meta[i], b.desc.tri_meta[i]
        assert (ma.edge_ab, ma.edge_bc, ma.edge_ca, ma.shape_idx) == (mb.edge_ab, mb.edge_bc, mb.edge_ca, mb.shape_idx)


def test_double_slit_scene_matches_reference_geometry():
    """scenes/diffraction_simple/bits/geometry.xml: wall + floor + 3 screen rectangles = 10 triangles; slit centres at +-W/2."""
    b = scenes.double_slits(res=64, spp=1).build()
    d = b.desc
    assert d.n_tris == 10 and d.n_shapes == 5 and d.n_emitters == 3
    xs = sorted({round(v * 1e3, 4) for i in range(d.n_tris) for v in (d.tris[i].ax, d.tris[i].bx, d.tris[i].cx) if abs(d.tris[i].az + 0.015) < 1e-6})
    assert xs == [-6.0, -0.5, -0.15, 0.15, 0.5, 6.0]


def test_etoile_like_scene_builds_and_itu_iors():
    """The etoile restatement: 563 shapes (ground + 562 boxes), ITU complex IORs at 10 GHz against hand values of
    sqrt(eps_r - i sigma/(eps0 omega)) (ITU-R P.2040-2 table 3 as in src/spectrum/util/spectrum_from_ITU.cpp:78-170)."""
    from wave_tracer_b200.scene import ITU, wavelen_to_wavenum
    b = scenes.etoile_like(res=64, spp=1).build()
    d = b.desc
    assert d.n_shapes == 563 and d.n_tris == 2 + 562 * 12 and d.n_emitters == 1 and (b.width, b.height) == (64, 48)
    assert d.integrator.type == 0 and d.integrator.direction == 1 and d.integrator.russian_roulette == 0 and d.integrator.max_depth == 16
    k = np.array([wavelen_to_wavenum(2.99792458e8 / 10e9)])
    for name, (eps, sig) in {"concrete": (5.24, 0.0462 * 10 ** 0.7822), "marble": (7.074, 0.0055 * 10 ** 0.9262), "wood": (1.99, 0.0047 * 10 ** 1.0718)}.items():
        want = np.sqrt(eps - 1j * sig / (8.8541878128e-12 * 2 * math.pi * 10e9))
        assert abs(ITU(name).value(k)[0] - want) < 1e-6 * abs(want)
    n = ITU("metal").value(k)[0]
    assert n.real > 1e3 and abs(n.real + n.imag) / n.real < 1e-3          # good conductor: n ~ (1 - i) sqrt(sigma / (2 eps0 omega))
