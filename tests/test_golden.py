"""Golden fixtures (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py from the CPU oracle -- the reference holds none and
cannot run here).  CPU: the oracle still reproduces them (guards the checker against drift).  GPU: the CUDA path matches the
frozen fixtures within the f32 tolerances of tests/test_gpu_parity.py."""
import ctypes as C
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden
import _oracle

G = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))
CASES = make_golden.cases()
COUNTERS = ("samples", "segments", "surface", "fsd", "null_", "splats")


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden_films(name):
    mk, spp = CASES[name]
    b = mk().build()
    blk, lgt, st = _oracle.render(b, spp=spp, seed=0x5EED, threads=1)
    assert [st[k] for k in COUNTERS] == list(G[name + "/counters"])
    # single-threaded oracle: f64 film sums in a fixed order -> exact up to libm differences between builds
    assert np.allclose(blk, G[name + "/block"], rtol=1e-9, atol=0) and np.allclose(lgt, G[name + "/light"], rtol=1e-9, atol=0)


def test_oracle_reproduces_golden_sobol_and_philox():
    from wave_tracer_b200 import sobol
    t = sobol.default_table()
    assert np.array_equal(np.array([[d, sj, aj] + list(mk) + [0] * (5 - len(mk)) for d, sj, aj, mk in t], np.int64), G["sobol/table"])
    n = 81
    num = np.zeros(n * 47, np.uint32); val = np.zeros(n * 47, np.float32)
    _oracle.lib().oracle_sobol_batch(sobol.to_abi(t), 0x5EED, 0, n, num.ctypes.data_as(C.POINTER(C.c_uint32)), val.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(num.reshape(n, 47), G["sobol/numerators_seed5EED_batch0_81pts"])
    rng = np.zeros(64, np.float32)
    _oracle.lib().oracle_rng(0x5EED, 123, 7, 64, rng.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(rng, G["rng/philox_seed5EED_pixel123_sample7"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_matches_golden_films(name):
    from wave_tracer_b200 import render
    mk, spp = CASES[name]
    b = mk().build()
    blk, lgt, st = render(b, spp=spp, seed=0x5EED)
    gold_c = dict(zip(COUNTERS, G[name + "/counters"]))
    assert st["samples"] == gold_c["samples"]
    for g, o in ((blk.astype(np.float64), G[name + "/block"]), (lgt.astype(np.float64), G[name + "/light"])):
        den = np.linalg.norm(o)
        if den > 0: assert np.linalg.norm(g - o) / den <= 1e-3, (name, np.linalg.norm(g - o) / den)
        else: assert np.abs(g).max() == 0


@pytest.mark.gpu
def test_gpu_matches_golden_sobol_and_philox():
    from wave_tracer_b200 import scenes, GpuScene, _abi as A
    from wave_tracer_b200.scene import Sobolld
    sc = scenes.cornell_like(res=16, spp=4, n_sphere=4); sc.sampler = Sobolld()
    gs = GpuScene(sc.build(), 0)
    n = 81; num = np.zeros(n * 47, np.uint32); val = np.zeros(n * 47, np.float32)
    A.check(A.lib().wtgpu_debug_sobol(gs.handle, 0x5EED, 0, n, num.ctypes.data_as(C.POINTER(C.c_uint32)), val.ctypes.data_as(C.POINTER(C.c_float))), "wtgpu_debug_sobol")
    assert np.array_equal(num.reshape(n, 47), G["sobol/numerators_seed5EED_batch0_81pts"])
    rng = np.zeros(64, np.float32)
    A.check(A.lib().wtgpu_debug_rng(0x5EED, 123, 7, 64, rng.ctypes.data_as(C.POINTER(C.c_float)), 0), "wtgpu_debug_rng")
    assert np.array_equal(rng, G["rng/philox_seed5EED_pixel123_sample7"])
    gs.close()
