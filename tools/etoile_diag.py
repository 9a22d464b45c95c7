#!/usr/bin/env python
"""Diagnostic: where does the etoile-like GPU film differ from the oracle's?  (run on the GPU box)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from wave_tracer_b200 import scenes, render, GpuScene
import _oracle
b = scenes.etoile_like(res=96, spp=4).build()
gs = GpuScene(b, 0)
for seed in (0x5EED, 1, 2, 3):
    for flags in (0, 8):
        blk, lgt, st = render(b, spp=4, seed=seed, gpu_scene=gs, flags=flags)
        oblk, olgt, ost = _oracle.render(b, spp=4, seed=seed)
        d = lgt.astype(np.float64) - olgt
        l2 = np.linalg.norm(d) / np.linalg.norm(olgt)
        big = np.argwhere(np.abs(d) > 1e-4 * np.abs(olgt).max())
        print(f"seed {seed:#x} flags {flags}: overflows {st['capacity_overflows']} segments {st['segments']}/{ost['segments']} fsd {st['fsd_interactions']}/{ost['fsd']} surf {st['surface_interactions']}/{ost['surface']} "
              f"splats {st['splats']}/{ost['splats']} rel-L2 {l2:.3e} flux {abs(d.sum()) / olgt.sum():.3e} max|o| {np.abs(olgt).max():.3e} |o|2 {np.linalg.norm(olgt):.3e} pixels differing {len(big)}")
        for y, x, c in big[:6]:
            print(f"      ({x},{y}) gpu {lgt[y, x, c]:.6e} oracle {olgt[y, x, c]:.6e}")
