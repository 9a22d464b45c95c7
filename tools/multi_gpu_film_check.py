#!/usr/bin/env python
"""Hardware check of the multi-GPU scheme (run under torchrun, N >= 2 GPUs): the film N ranks render by sample partition and reduce over NCCL
(wave_tracer_b200.parallel.render_distributed) equals the film one rank renders alone, up to the order of the f32 additions; the developed image
is produced on the device on rank 0 (wtgpu_develop_device).   torchrun --nproc-per-node N tools/multi_gpu_film_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from wave_tracer_b200 import scenes, GpuScene, render, develop
from wave_tracer_b200.parallel import render_distributed, develop_on_device
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local)); torch.cuda.set_device(local)
ok = True
for name, mk, spp in (("cornell_box plt_bdpt rgb", lambda: scenes.cornell_box(res=64, spp=8, dragon_tris=11520, bunny_tris=5120, lut=(512, 256)), 8),
                      ("etoile plt_path utd", lambda: scenes.etoile_like(res=96, spp=8, detail=3), 8)):
    b = mk().build(table_size=256)
    gs = GpuScene(b, local)
    blk, lgt, st = render_distributed(gs, spp)
    if rank == 0:
        img_n = develop_on_device(gs, spp, blk, lgt).cpu().numpy().astype(np.float64)
        b1, l1, st1 = render(b, spp=spp, gpu_scene=gs)
        img_1 = develop(b, spp, b1, l1)
        rel = float(np.linalg.norm(img_n - img_1) / max(np.linalg.norm(img_1), 1e-300))
        print("%s: %d ranks vs 1 rank: rel-L2 %.3e (samples %d on rank 0 of %d)" % (name, world, rel, st["samples"], st1["samples"]), flush=True)
        ok = ok and rel < 1e-5
    gs.close()
    dist.barrier()
if rank == 0: print("MULTI_GPU_FILM_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
