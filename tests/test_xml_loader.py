"""Scene XML front-end (wave_tracer_b200/xml_loader.py): files in the reference's scene format load into the same tables as the equivalent
Python-API scene.  The fixture tests/data/slit_bench.xml was written for this repository; when the reference tree is mounted
(/root/reference, this container only) the reference's own double_slits.xml is loaded too and compared with scenes.double_slits()."""
import ctypes as C
import math
import os
import numpy as np
import pytest

from wave_tracer_b200 import _abi as A, scenes, xml_loader
from wave_tracer_b200 import (Scene, PltPath, Film, VirtualPlane, Spot, Discrete, Diffuse, SurfaceSPM, Gaussian, TwoSided, Composite, Binned, rectangle, lookat)
import _oracle

HERE = os.path.dirname(os.path.abspath(__file__))
MM = 1e-3


def _raw(p, n, t):
    return bytes(C.cast(p, C.POINTER(C.c_char * (C.sizeof(t) * n))).contents) if n else b""


def _same_tables(b1, b2):
    d1, d2 = b1.desc, b2.desc
    for f in ("n_nodes", "n_leaves", "n_tris", "n_edges", "n_shapes", "n_spectra", "n_bsdfs", "n_emitters", "n_kdist_data"):
        assert getattr(d1, f) == getattr(d2, f), f
    for name, t, n in (("tris", A.Tri, d1.n_tris), ("nodes", A.Node, d1.n_nodes), ("edges", A.Edge, d1.n_edges), ("emitters", A.Emitter, d1.n_emitters),
                       ("bsdfs", A.Bsdf, d1.n_bsdfs), ("spectra", A.Spectrum, d1.n_spectra)):
        assert _raw(getattr(d1, name), n, t) == _raw(getattr(d2, name), n, t), name
    assert bytes(d1.sensor) == bytes(d2.sensor) and bytes(d1.integrator) == bytes(d2.integrator)


def _slit_bench_api(res=96, spp=4, lam_mm=.08, gap=.9, Zs=-12.0, with_floor=True):
    lam = lam_mm * MM
    sc = Scene()
    sc.integrator = PltPath(max_depth=12, direction="forward", russian_roulette=False)
    film = Film(res, res // 3, [Discrete(lam)], rfilter_scale=.1)
    sc.sensor = VirtualPlane(lookat((0, 0, (40 - .001) * MM), (0, 0, 2 * MM), (0, -1, 0)), (200 * MM, 200 / 3 * MM), film, alpha=math.radians(.002), samples=spp)
    sc.add_emitter(Spot(lookat((0, 0, -400 * MM), (0, 0, 0), (1, 0, 0)), Discrete(lam, 900.0), cutoff_angle=math.radians(.3), beam_width=math.radians(.15)))
    mat_screen = TwoSided(SurfaceSPM(IOR=complex(1, 80), profile=Gaussian(sigma=50.0)))
    mat_floor = TwoSided(Composite([(1e-6, 1.0, Diffuse(.15))]))
    mat_wall = TwoSided(Diffuse(Binned([(300e-9, 800e-9, .5), (1e-6, 1.0, .85)])))
    def rect(p, x, y, m): sc.add_shape(rectangle(np.array(p) * MM, np.array(x) * MM, np.array(y) * MM), m)
    rect((-80, -15, 40), (160, 0, 0), (0, 30, 0), mat_wall)
    if with_floor: rect((-80, -15, -450), (160, 0, 0), (0, 0, 490), mat_floor)
    rect((-6, -15, Zs), (6 - gap / 2, 0, 0), (0, 30, 0), mat_screen)
    rect((gap / 2, -15, Zs), (6 - gap / 2, 0, 0), (0, 30, 0), mat_screen)
    return sc


def test_expressions_and_quantities():
    q = xml_loader.quantity
    assert q("(50-.0001) mm", "len") == pytest.approx(49.9999e-3) and q(".001°", "ang") == pytest.approx(math.radians(.001)) and q("0mm", "len") == 0
    assert q("10GHz", "wavelength") == pytest.approx(2.99792458e8 / 1e10) and q("550 nm", "wavelength") == pytest.approx(550e-9)
    assert xml_loader.integer("1440/4") == 360 and xml_loader.boolean("(true==true && false==false)") and not xml_loader.boolean("(1 && 0>0)")
    assert xml_loader.complex_value("(1,100i)") == complex(1, 100) and xml_loader.complex_value("1.5") == 1.5
    assert xml_loader.qvec("0mm, 0mm, (2*3) mm", "len", 3) == [0, 0, pytest.approx(6e-3)]
    assert xml_loader.qrange("300nm .. 800nm", "wavelength") == (pytest.approx(300e-9), pytest.approx(800e-9))
    with pytest.raises(xml_loader.SceneXmlError): q("__import__('os')", None)
    with pytest.raises(xml_loader.SceneXmlError): q("5", "len")
    assert xml_loader.parse_defines("res=1440,spp=1024") == {"res": "1440", "spp": "1024"}


@pytest.mark.parametrize("defines,kw", [({}, {}), ({"res": "48", "gap": "1.3", "with_floor": "false"}, dict(res=48, gap=1.3, with_floor=False))])
def test_fixture_xml_equals_python_api_scene(defines, kw):
    b1 = xml_loader.load_scene(os.path.join(HERE, "data", "slit_bench.xml"), defines).build()
    b2 = _slit_bench_api(**kw).build()
    _same_tables(b1, b2)
    o1 = _oracle.render(b1, spp=2, threads=1); o2 = _oracle.render(b2, spp=2, threads=1)
    assert o1[1].sum() > 0 and np.array_equal(o1[1], o2[1]) and np.array_equal(o1[0], o2[0])


def test_unsupported_elements_fail_loudly(tmp_path):
    p = tmp_path / "bad.xml"
    p.write_text('<scene version="0.1.0"><integrator type="plt_path"/><sensor type="fisheye"><film type="array"><response type="monochromatic">'
                 '<spectrum type="discrete" wavelength="1mm"/></response></film></sensor></scene>')
    with pytest.raises(xml_loader.SceneXmlError, match="fisheye"):
        xml_loader.load_scene(str(p))
    p.write_text('<scene version="0.1.0"><integrator type="plt_path"/><volume/></scene>')
    with pytest.raises(xml_loader.SceneXmlError):
        xml_loader.load_scene(str(p))


REF = "/root/reference/scenes/diffraction_simple/double_slits.xml"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted (GPU box)")
def test_reference_double_slits_xml_equals_procedural_restatement():
    """BASELINE configs[1], from the untouched reference file (+ its bits/geometry.xml include), against scenes.double_slits(): the same
    triangles, BVH, edges, emitters, sensor and integrator tables, and bit-identical oracle films.  (The composite floor bsdf keeps only
    the bin the 0.05 mm sensor can reach, so bsdf/spectrum tables are compared through the films.)"""
    b1 = xml_loader.load_scene(REF, {"res": "64", "spp": "4"}, lut=(256, 128)).build()
    b2 = scenes.double_slits(res=64, spp=4, integrator="plt_bdpt", lut=(256, 128)).build()
    d1, d2 = b1.desc, b2.desc
    for name, t, n in (("tris", A.Tri, d1.n_tris), ("nodes", A.Node, d1.n_nodes), ("edges", A.Edge, d1.n_edges), ("emitters", A.Emitter, d1.n_emitters)):
        assert _raw(getattr(d1, name), n, t) == _raw(getattr(d2, name), n, t), name
    assert bytes(d1.sensor) == bytes(d2.sensor) and bytes(d1.integrator) == bytes(d2.integrator)
    o1 = _oracle.render(b1, spp=4, threads=1); o2 = _oracle.render(b2, spp=4, threads=1)
    assert o1[1].sum() > 0 and np.array_equal(o1[0], o2[0]) and np.array_equal(o1[1], o2[1])


REF_ETOILE = "/root/reference/scenes/sionna_etoile/etoile.xml"


@pytest.mark.skipif(not os.path.exists(REF_ETOILE), reason="reference tree not mounted (GPU box)")
def test_reference_etoile_xml_matches_the_restatement_outside_its_meshes():
    """BASELINE configs[3]: every shape of etoile.xml is a PLY mesh (Git-LFS stubs) -- missing_meshes="skip" lists all 563 and loads the rest.
    Integrator, coverage sensor, the 10 GHz point emitter and the five ITU materials equal what scenes.etoile_like() restates; the file's three
    optical emitters (D65 / D55 illuminants) carry no power at the sensor's wavenumber."""
    from wave_tracer_b200 import TwoSided, Composite, SurfaceSPM, ITU, Diffuse, Const, Point, Directional
    sc = xml_loader.load_scene(REF_ETOILE, {"res": "720", "spp": "1024", "wavelength": "10GHz"}, missing_meshes="skip")
    ref = scenes.etoile_like(res=720, spp=1024)
    assert len(sc.skipped_shapes) == 563 and sc.skipped_shapes[0] == ("mesh-Plane", "meshes/Plane.ply") and len(sc.shapes) == 0
    assert vars(sc.integrator) == vars(ref.integrator)
    s, r = sc.sensor, ref.sensor
    assert type(s) is type(r) and np.array_equal(s.to_world, r.to_world) and tuple(s.extent) == tuple(r.extent) and s.alpha == r.alpha and s.samples == r.samples and s.rt == r.rt
    assert (s.film.width, s.film.height, s.film.rfilter_scale) == (r.film.width, r.film.height, r.film.rfilter_scale) and s.film.response[0].lines == r.film.response[0].lines
    e, q = sc.emitters[0], ref.emitters[0]
    assert isinstance(e, Point) and tuple(e.position) == tuple(q.position) and e.spectrum.lines == q.spectrum.lines and e.pse == q.pse
    assert len(sc.emitters) == 4 and isinstance(sc.emitters[2], Directional)
    k = np.array([s.film.response[0].lines[0][0]], np.float64)
    for extra in sc.emitters[1:]:
        assert np.all(extra.spectrum.value(k) == 0)
    for m in ("marble", "metal", "brick", "wood", "concrete"):
        b = sc.xml_bsdfs["mat-itu_" + m]
        assert isinstance(b, TwoSided) and isinstance(b.nested, Composite) and len(b.nested.bins) == 2
        lo, hi, radio = b.nested.bins[1]
        assert (lo, hi) == (pytest.approx(.1e-3), pytest.approx(1.0)) and isinstance(radio, SurfaceSPM) and isinstance(radio.IOR, ITU) and radio.IOR.params == ITU.TABLE[m]
        assert isinstance(radio.ts, Const) and radio.ts.v == 0
    with pytest.raises(xml_loader.SceneXmlError, match="missing_meshes"):
        xml_loader.load_scene(REF_ETOILE, {"wavelength": "10GHz"})


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted (GPU box)")
def test_reference_double_slits_and_reflectors_xml_loads():
    """The plt_path-forward sibling of BASELINE configs[1] (what `bench.py --integrator plt_path` restates) loads and renders on the oracle."""
    sc = xml_loader.load_scene(os.path.join(os.path.dirname(REF), "double_slits_and_reflectors.xml"), {"res": "64", "spp": "2"})
    assert type(sc.integrator).__name__ == "PltPath" and sc.integrator.direction == "forward"
    b = sc.build()
    o = _oracle.render(b, spp=2, threads=1)
    assert o[2]["samples"] == 64 * 16 * 2 and o[1].sum() > 0
