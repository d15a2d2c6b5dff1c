"""Multi-GPU plumbing for one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Rendering shards naturally: the scene is replicated, every rank renders its share
of the frame through the C ABI (`rank`/`world_size`/`split` of pbrs_render_opts).  There is no
exchange on the data path other than the film:
  tiles    64x64 tiles, tile t -> rank t % N.  Films are disjoint.
           device film: one NCCL reduce(sum) to rank 0, bit-identical to the single-GPU film (x + 0);
           host film:   every rank copies ITS OWN tiles straight into one host film that all ranks
                        map (`SharedHostFilm`, PBRS_FLAG_OWN_TILES_ONLY): no inter-GPU traffic at all
  samples  sample i of every pixel -> rank i % N; ranks return raw partial sums
           (PBRS_FLAG_RAW_SUM), one NCCL reduce(sum) to rank 0, which scales by 1/spp
           (src/main.rs:208) and, for a host film, copies it out once
(One process driving N GPUs needs none of this: pbrs_render with num_gpus = N, include/pbrs_gpu.h.)
"""
import mmap
import os

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


class SharedHostFilm:
    """One [H, W, 3] float32 host film mapped by every rank of the job (POSIX shared memory), page-
    locked in each process when a CUDA library is given, so that each rank's device can DMA its own
    tiles straight into it.  Rank 0 creates the segment; collective: every rank constructs it."""

    def __init__(self, height, width, api=None, name=None):
        world, rank = world_rank()
        self.api = api
        self.bytes = int(height) * int(width) * 12
        if name is None:
            tag = [f"pbrs_film_{os.getpid()}_{np.random.default_rng().integers(1 << 30)}" if rank == 0 else None]
            if world > 1:
                dist.broadcast_object_list(tag, src=0)
            name = tag[0]
        self.path = os.path.join("/dev/shm", name)
        self.owner = rank == 0
        if self.owner:
            with open(self.path, "wb") as f:
                f.truncate(self.bytes)
        if world > 1:
            dist.barrier()
        self.fd = os.open(self.path, os.O_RDWR)
        self.map = mmap.mmap(self.fd, self.bytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        self.array = np.frombuffer(self.map, dtype=np.float32).reshape(int(height), int(width), 3)
        self.registered = False
        if api is not None:
            self.registered = api["host_register"](self.array.ctypes.data, self.bytes) == 0

    def close(self):
        world, _ = world_rank()
        if self.registered:
            self.api["host_unregister"](self.array.ctypes.data)
            self.registered = False
        self.array = None
        try:
            self.map.close()
        except BufferError:
            pass
        os.close(self.fd)
        if world > 1:
            dist.barrier()
        if self.owner and os.path.exists(self.path):
            os.unlink(self.path)


def render_sharded(handle, integrator, msaa, max_depth=5, split="tiles", seed=0x5EED, paths_in_flight=0, device_film=None, stream_ptr=None,
                   host_film=None):
    """Every rank calls this; on rank 0 the returned film holds the whole frame.

    device_film: a CUDA float32 tensor [H, W, 3] -> the film never leaves the GPUs
                 (pbrs_render_device + NCCL reduce).
    host_film:   a SharedHostFilm -> the end-to-end path.  Tile split: every rank's pbrs_render copies
                 its own tiles into the shared film, then one barrier.  Sample split: NCCL reduce of
                 the device films, rank 0 copies the result out (needs device_film as scratch).
    neither:     pbrs_render into a private numpy film, reduced through torch.distributed on
                 whatever backend the process group has (gloo on CPU-only test runs)."""
    world, rank = world_rank()
    kw = shard_kwargs(split)
    raw_spp = msaa * msaa if (world > 1 and split == "samples") else None
    common = dict(integrator=integrator, msaa=msaa, max_depth=max_depth, seed=seed, paths_in_flight=paths_in_flight, **kw)
    if host_film is not None and (world == 1 or split == "tiles"):
        if world > 1:
            common["flags"] |= K.FLAG_OWN_TILES_ONLY
        handle.render(want_stats=False, out=host_film.array, **common)
        if world > 1:
            dist.barrier()
        return host_film.array
    if device_film is not None:
        if stream_ptr is None:  # the stream the reduce below is ordered against
            stream_ptr = torch.cuda.current_stream(device_film.device).cuda_stream
        handle.render_device(device_film.data_ptr(), stream=stream_ptr, **common)
        film_reduce(device_film, raw_spp)
        if host_film is not None:
            if rank == 0:
                torch.from_numpy(host_film.array).copy_(device_film)  # page-locked target: one DMA
            else:
                torch.cuda.current_stream(device_film.device).synchronize()
            return host_film.array
        return device_film
    film, _ = handle.render(want_stats=False, **common)
    t = torch.from_numpy(film)
    if world > 1 and dist.get_backend() == "nccl":
        t = t.cuda()
    t = film_reduce(t, raw_spp)
    return t.cpu().numpy() if t.is_cuda else film
