"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Rendering shards naturally: the scene is replicated, every rank renders its share
of the frame through the C ABI (`rank`/`world_size`/`split` of pbrs_render_opts) and the partial
films are summed by ONE reduce to rank 0.  There is no other exchange on this path.
  tiles    64x64 tiles, tile t -> rank t % N; films are disjoint, so the sum is bit-identical to
           the single-GPU film (x + 0)
  samples  sample i of every pixel -> rank i % N; ranks return raw partial sums
           (PBRS_FLAG_RAW_SUM) and rank 0 scales the reduced sum by 1/spp (src/main.rs:208)
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi as K


def split_for(workload):
    """C5 (1024 spp over 8 GPUs) splits samples, everything else splits tiles (BASELINE.json configs)."""
    return "samples" if workload == "c5" else "tiles"


def world_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_kwargs(split, flags=0):
    """rank / world_size / split / flags of pbrs_render_opts for this process."""
    world, rank = world_rank()
    raw = world > 1 and split == "samples"
    return dict(rank=rank, world_size=world, split=split, flags=flags | (K.FLAG_RAW_SUM if raw else 0))


def film_reduce(film, spp_if_raw_sum=None, dst=0):
    """Sum of the ranks' films on `dst` (a torch tensor, reduced in place; NCCL for CUDA tensors,
    gloo for CPU tensors)."""
    world, rank = world_rank()
    if world > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    if spp_if_raw_sum and (world == 1 or rank == dst):
        film.mul_(1.0 / float(spp_if_raw_sum))
    return film


def render_sharded(handle, integrator, msaa, max_depth=5, split="tiles", seed=0x5EED, paths_in_flight=0, device_film=None, stream_ptr=None):
    """Every rank calls this; on rank 0 the returned film holds the whole frame.

    device_film: a CUDA float32 tensor [H, W, 3] -> the film never leaves the GPUs
    (pbrs_render_device + NCCL reduce).  Otherwise the host path: pbrs_render into a numpy film,
    reduced through torch.distributed on whatever backend the process group has."""
    world, _ = world_rank()
    kw = shard_kwargs(split)
    raw_spp = msaa * msaa if (world > 1 and split == "samples") else None
    common = dict(integrator=integrator, msaa=msaa, max_depth=max_depth, seed=seed, paths_in_flight=paths_in_flight, **kw)
    if device_film is not None:
        handle.render_device(device_film.data_ptr(), stream=stream_ptr, **common)
        return film_reduce(device_film, raw_spp)
    film, _ = handle.render(want_stats=False, **common)
    t = torch.from_numpy(film)
    if world > 1 and dist.get_backend() == "nccl":
        t = t.cuda()
    t = film_reduce(t, raw_spp)
    return t.cpu().numpy() if t.is_cuda else film
