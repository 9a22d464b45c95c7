"""Multi-GPU driver: one process per GPU (torch.distributed), samples partitioned across ranks,
every rank accumulates into its own full-size film, ONE reduce(sum) of the film at the end -- the GPU analogue of the
reference's per-worker light images summed at develop time (film_storage.hpp:155-158, 276-288).
There is no per-bounce communication, hence no fused compute+collective kernel on this path."""
import numpy as np


def partition_samples(spp, rank, world):
    """Contiguous, disjoint sample ranges; rank r gets [begin, end).  Union over ranks == [0, spp)."""
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_films(block, light, dst=0):
    """Sum the per-rank films onto rank `dst` (torch tensors on any backend).  A single collective on one flat buffer."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return block, light
    flat = torch.cat([block.reshape(-1), light.reshape(-1)])
    dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM)
    nb = block.numel()
    return flat[:nb].reshape(block.shape), flat[nb:].reshape(light.shape)


def render_distributed(gpu_scene, spp, seed=0x5EED, pool_size=0, flags=0):
    """Renders this rank's sample range on its GPU into device-resident torch tensors and reduces to rank 0.
    Returns (film_block, film_light, stats) -- films are valid on rank 0."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    b = gpu_scene.built
    dev = torch.device("cuda", gpu_scene.device)
    block = torch.zeros((b.height, b.width, b.channels, 2), dtype=torch.float32, device=dev)
    light = torch.zeros((b.height, b.width, b.channels), dtype=torch.float32, device=dev)
    s0, s1 = partition_samples(spp, rank, world)
    stream = torch.cuda.current_stream(dev).cuda_stream
    # A rank must never leave for an exception while the others wait in the collective: the status is agreed on first.
    st, err = None, None
    try:
        st = gpu_scene.render_into(block.data_ptr(), light.data_ptr(), spp, seed, (s0, s1), None, True, pool_size, flags, stream)
    except Exception as e:      # noqa: BLE001 -- reported on every rank below
        err = e
    if world > 1:
        bad = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()) and err is None:
            err = RuntimeError("wtgpu_render failed on another rank")
    if err is not None:
        raise err
    block, light = reduce_films(block, light)
    return block, light, st


def develop_on_device(gpu_scene, spp, block, light):
    """film develop (film_storage.hpp:256-291, 354-358) of device-resident films on their GPU (wtgpu_develop_device): returns a torch tensor
    [H][W][C] -- on rank 0 after the reduce, so that only the developed image crosses PCIe."""
    import ctypes as C
    import torch
    from . import _abi as A
    b = gpu_scene.built
    out = torch.empty((b.height, b.width, b.channels), dtype=torch.float32, device=block.device)
    stream = torch.cuda.current_stream(block.device).cuda_stream
    A.check(A.lib().wtgpu_develop_device(C.byref(b.desc.sensor), spp, block.data_ptr(), light.data_ptr(), out.data_ptr(), stream, gpu_scene.device), "wtgpu_develop_device")
    return out
