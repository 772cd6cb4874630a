"""Host-side logic of the temporal shard, on CPU: the frame partition the engine uses and the
post-sampling exchange of latent slabs over torch.distributed (gloo, world_size 2 and 3)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kandinsky.models.parallelize import frame_partition, gather_frames, parallelize_dit


@pytest.mark.parametrize("T,world", [(31, 1), (31, 2), (31, 4), (31, 8), (61, 8), (8, 8), (3, 2), (241, 8)])
def test_frame_partition_covers_every_frame_once_and_is_balanced(T, world):
    parts = frame_partition(T, world)
    assert len(parts) == world
    assert parts[0][0] == 0 and sum(n for _, n in parts) == T
    for (f0, n), (g0, _) in zip(parts, parts[1:]):
        assert f0 + n == g0                                   # contiguous, ordered
    counts = [n for _, n in parts]
    assert max(counts) - min(counts) <= 1 and min(counts) >= 1
    assert counts == sorted(counts, reverse=True)             # the longer slabs come first (engine_set_grid)


def test_frame_partition_rejects_more_ranks_than_frames():
    with pytest.raises(ValueError):
        frame_partition(3, 4)
    with pytest.raises(ValueError):
        frame_partition(3, 0)


class _FakeDit:
    """Stands in for the engine-backed module: records what parallelize_dit hands to k5_dist_init."""

    def __init__(self, world, rank=0):
        self.dist_world, self.rank, self.got = world, rank, None

    def dist_export(self):
        return b"handle-of-rank-%d" % self.rank

    def dist_init(self, rank, world, handles):
        self.got = (rank, world, list(handles))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(T * 2 * 3 * 4, dtype=torch.float32).reshape(T, 2, 3, 4)
        mine = torch.full_like(full, -1.0)                    # other ranks' frames are stale
        f0, n = frame_partition(T, world)[rank]
        mine[f0:f0 + n] = full[f0:f0 + n]
        out = gather_frames(_FakeDit(world), mine)
        ret[rank] = bool(torch.equal(out, full))
        fake = parallelize_dit(_FakeDit(world, rank))         # handle exchange: every rank sees all, in rank order
        assert fake.got == (rank, world, [b"handle-of-rank-%d" % r for r in range(world)])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 31), (3, 7)])
def test_gather_frames_over_gloo(world, T):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), T, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
