"""examples/host_render.cpp: a C++ host that drives the hot path through the C-ABI alone (no Python, no torch) -- the shape of the glue
INTEGRATION.md describes for wave_tracer's render driver.  CPU: the headers are plain C, the example compiles, links against the product
library and refuses to run without a GPU.  GPU: its film equals the film of the same scene built through the Python layer."""
import math
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "host_render")
MM = 1e-3


def _build_example():
    import __graft_entry__ as g
    g.build_example()
    assert os.path.exists(EXE)


def test_headers_are_plain_c_and_cxx():
    for cmd in (["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c"], ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++"]):
        for h in ("wtgpu.h", "wthost.h"):
            r = subprocess.run(cmd + [os.path.join(ROOT, "include", h)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr


def test_cxx_host_example_builds_and_fails_loudly_without_gpu():
    _build_example()
    r = subprocess.run([EXE, "32", "2"], capture_output=True, text=True, timeout=300)
    from wave_tracer_b200 import _abi as A
    if A.lib().wtgpu_device_count() > 0:
        assert r.returncode == 0 and "host_render:" in r.stdout, (r.returncode, r.stderr)
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stdout, r.stderr)


def _python_twin(res, spp):
    from wave_tracer_b200 import Scene, PltPath, Film, VirtualPlane, Spot, Discrete, Diffuse, TwoSided, rectangle, lookat
    lam = .08 * MM
    sc = Scene()
    sc.integrator = PltPath(max_depth=12, direction="forward", russian_roulette=False, fsd=True)
    film = Film(res, res // 3, [Discrete(lam)], rfilter_scale=.1)
    sc.sensor = VirtualPlane(lookat((0, 0, (40 - .001) * MM), (0, 0, 2 * MM), (0, -1, 0)), (200 * MM, 200 / 3 * MM), film, alpha=math.radians(.002), samples=spp)
    sc.add_emitter(Spot(lookat((0, 0, -400 * MM), (0, 0, 0), (1, 0, 0)), Discrete(lam, 900.0), cutoff_angle=math.radians(.3), beam_width=math.radians(.15)))
    wall, screen = TwoSided(Diffuse(.85)), TwoSided(Diffuse(.3))
    def rect(p, x, y, m): sc.add_shape(rectangle(np.array(p) * MM, np.array(x) * MM, np.array(y) * MM), m)
    rect((-80, -15, 40), (160, 0, 0), (0, 30, 0), wall)
    rect((-6, -15, -12), (6 - .45, 0, 0), (0, 30, 0), screen)
    rect((.45, -15, -12), (6 - .45, 0, 0), (0, 30, 0), screen)
    return sc


@pytest.mark.gpu
def test_cxx_host_example_film_equals_python_layer(tmp_path):
    from wave_tracer_b200 import render, develop
    _build_example()
    res, spp = 192, 8
    out = str(tmp_path / "img.f32")
    r = subprocess.run([EXE, str(res), str(spp), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    img_c = np.fromfile(out, np.float32).reshape(res // 3, res, 1)
    b = _python_twin(res, spp).build()
    blk, lgt, st = render(b, spp=spp, seed=0x5EED)
    img_p = develop(b, spp, blk, lgt)
    print(r.stdout.strip(), "| python sum %.9e" % img_p.sum())
    assert img_p.sum() > 0 and st["samples"] == res * (res // 3) * spp
    num = np.linalg.norm(img_c.astype(np.float64) - img_p); den = np.linalg.norm(img_p)
    assert num <= 1e-5 * den, (num, den)       # same tables, same streams: only the order of the f32 film atomics differs
