"""The N > 1 path on CPU: two processes, gloo backend, the sharding logic of pbrs_b200/dist.py
(rank/world/split -> pbrs_render_opts, one film reduce to rank 0).  The renderer underneath is
the host build of the product's stage functions (tests/hostsim), since no GPU exists here."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, split, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from pbrs_b200 import dist as pdist
    from tests import hostsim
    from tests.util import SMALL_SCENES

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h = SMALL_SCENES["cornell"]().realize(hostsim.load())
    film = pdist.render_sharded(h, "path", msaa=2, max_depth=3, split=split)
    if rank == 0:
        np.save(out_path, film)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("split", ["tiles", "samples"])
def test_two_ranks_reduce_to_the_single_rank_film(tmp_path, hostsim_api, split):
    from tests.util import SMALL_SCENES, bits_equal
    out = str(tmp_path / "film.npy")
    port = 29500 + (os.getpid() % 2000) + (0 if split == "tiles" else 1)
    mp.spawn(_worker, args=(2, port, split, out), nprocs=2, join=True)
    got = np.load(out)
    h = SMALL_SCENES["cornell"]().realize(hostsim_api)
    want, _ = h.render(integrator="path", msaa=2, max_depth=3)
    if split == "tiles":
        assert bits_equal(got, want).all()  # disjoint films: x + 0 is exact
    else:
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)  # fp32 summation order only


def test_shard_kwargs_single_process():
    from pbrs_b200 import dist as pdist
    kw = pdist.shard_kwargs("samples")
    assert kw == dict(rank=0, world_size=1, split="samples", flags=0)
    assert pdist.split_for("c5") == "samples" and pdist.split_for("c4") == "tiles"
