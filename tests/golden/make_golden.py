#!/usr/bin/env python
"""Regenerates the golden fixtures of tests/golden/ from the CPU oracle (oracle/liboracle.so).

The reference holds no golden vectors and cannot be compiled or imported here, so these fixtures pin the ORACLE against
drift (tests/test_golden.py, CPU) and give the GPU tests a second, frozen target: they are outputs of the oracle at the commit that made
them, on fixed seeds.  Run from the repo root:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))


def cases():
    from wave_tracer_b200 import scenes
    from wave_tracer_b200.scene import Sobolld
    def sob(sc): sc.sampler = Sobolld(); return sc
    return {
        "double_slits_path_fwd_utd_64x16_spp4": (lambda: scenes.double_slits(res=64, spp=4, with_directional=True), 4),
        "double_slits_bdpt_fraunhofer_64x16_spp4": (lambda: scenes.double_slits(res=64, spp=4, with_directional=False, integrator="plt_bdpt", lut=(256, 128)), 4),
        "cornell_path_bwd_24_spp4": (lambda: scenes.cornell_like(res=24, spp=4, n_sphere=6), 4),
        "cornell_bdpt_24_spp4": (lambda: scenes.cornell_like(res=24, spp=4, n_sphere=6, integrator="plt_bdpt"), 4),
        "cornell_path_bwd_sobolld_24_spp4": (lambda: sob(scenes.cornell_like(res=24, spp=4, n_sphere=6)), 4),
        "etoile_like_fwd_utd_48x36_spp2": (lambda: scenes.etoile_like(res=48, spp=2), 2),
    }


def main():
    import _oracle
    from wave_tracer_b200 import sobol
    out = {}
    for name, (mk, spp) in cases().items():
        b = mk().build()
        blk, lgt, st = _oracle.render(b, spp=spp, seed=0x5EED, threads=1)
        out[name + "/block"] = blk.astype(np.float64); out[name + "/light"] = lgt.astype(np.float64)
        out[name + "/counters"] = np.array([st[k] for k in ("samples", "segments", "surface", "fsd", "null_", "splats")], np.int64)
        print(name, st["samples"], st["segments"], float(blk[..., 0].sum()), float(lgt.sum()))
    t = sobol.default_table()
    n = 81
    num = np.zeros(n * 47, np.uint32); val = np.zeros(n * 47, np.float32)
    _oracle.lib().oracle_sobol_batch(sobol.to_abi(t), 0x5EED, 0, n, num.ctypes.data_as(C.POINTER(C.c_uint32)), val.ctypes.data_as(C.POINTER(C.c_float)))
    out["sobol/numerators_seed5EED_batch0_81pts"] = num.reshape(n, 47)
    out["sobol/table"] = np.array([[d, sj, aj] + list(mk) + [0] * (5 - len(mk)) for d, sj, aj, mk in t], np.int64)
    rng = np.zeros(64, np.float32)
    _oracle.lib().oracle_rng(0x5EED, 123, 7, 64, rng.ctypes.data_as(C.POINTER(C.c_float)))
    out["rng/philox_seed5EED_pixel123_sample7"] = rng
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_golden.npz"), os.path.getsize(os.path.join(HERE, "oracle_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
