"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Rendering shards naturally: the scene is replicated, every rank renders its share
of the frame through the C ABI (`rank`/`world_size`/`split` of pbrs_render_opts) and the partial
films are summed by ONE NCCL reduce to rank 0.  There is no other exchange on this path.
  tiles    64x64 tiles, tile t -> rank t % N; films are disjoint, so the sum is bit-identical to
           the single-GPU film (x + 0)
  samples  sample i of every pixel -> rank i % N; ranks return raw partial sums
           (PBRS_FLAG_RAW_SUM) and rank 0 scales the reduced sum by 1/spp (src/main.rs:208)
"""
import torch.distributed as dist

from . import _capi as K


def split_for(workload):
    """C5 (1024 spp over 8 GPUs) splits samples, everything else splits tiles (BASELINE.json configs)."""
    return "samples" if workload == "c5" else "tiles"


def film_reduce(film, spp_if_raw_sum=None, dst=0):
    """Sum of the ranks' films on `dst` (torch tensor on this rank's device, reduced in place)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
        if spp_if_raw_sum and dist.get_rank() == dst:
            film.mul_(1.0 / float(spp_if_raw_sum))
    elif spp_if_raw_sum:
        film.mul_(1.0 / float(spp_if_raw_sum))
    return film


def render_sharded(handle, film, stream_ptr, integrator, msaa, max_depth=5, split="tiles", seed=0x5EED, paths_in_flight=0):
    """Every rank calls this; rank 0's `film` ends up holding the whole frame."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    raw = world > 1 and split == "samples"
    handle.render_device(film.data_ptr(), stream=stream_ptr, integrator=integrator, msaa=msaa, max_depth=max_depth, seed=seed,
                         rank=rank, world_size=world, split=split, flags=(K.FLAG_RAW_SUM if raw else 0), paths_in_flight=paths_in_flight)
    return film_reduce(film, msaa * msaa if raw else None)
