"""N>1 host logic on CPU: world_size-2 gloo.  Each rank renders its sample range (the oracle stands in for the GPU kernel:
tests may use it), the films are summed with the same reduce the GPU path uses, and the result equals a single-rank render."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wave_tracer_b200.parallel import partition_samples, reduce_films


def test_partition_covers_range():
    for spp in (1, 7, 8, 1024):
        for world in (1, 2, 3, 8):
            rs = [partition_samples(spp, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == spp and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _oracle
    from wave_tracer_b200 import scenes
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = scenes.double_slits(res=64, spp=4, with_directional=False).build()
    blk, lgt, _ = _oracle.render(b, spp=4, sample_range=partition_samples(4, rank, world), threads=2)
    tb, tl = reduce_films(torch.from_numpy(blk), torch.from_numpy(lgt))
    if rank == 0:
        q.put((tb.numpy().copy(), tl.numpy().copy()))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_film_reduce_equals_single_rank():
    import _oracle
    from wave_tracer_b200 import scenes
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    blk2, lgt2 = q.get(timeout=300)
    for p in ps: p.join(timeout=60)
    b = scenes.double_slits(res=64, spp=4, with_directional=False).build()
    blk1, lgt1, _ = _oracle.render(b, spp=4, threads=2)
    assert lgt1.sum() > 0
    assert np.allclose(lgt2, lgt1, rtol=1e-12, atol=1e-12 * lgt1.max()) and np.allclose(blk2, blk1, rtol=1e-12)
