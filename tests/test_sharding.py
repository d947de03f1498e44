"""CPU tier: the N>1 path. Frame pairs are sharded across ranks with no data-path collective; this runs the sharding
logic under a real world_size-2 gloo process group (what bench.py does with NCCL on the GPU box) and checks coverage."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from model.sharding import frames_needed, shard_pairs


@pytest.mark.parametrize("frames,world", [(600, 1), (600, 2), (600, 4), (600, 8), (300, 8), (5, 8), (2, 2)])
def test_shards_partition_the_pairs(frames, world):
    seen = []
    for r in range(world):
        first, n = shard_pairs(frames, world, r)
        seen += list(range(first, first + n))
        fr = frames_needed(frames, world, r)
        assert (fr is None) == (n == 0)
        if fr:
            assert fr == (first, first + n) and fr[1] <= frames - 1
    assert seen == list(range(frames - 1))
    sizes = [shard_pairs(frames, world, r)[1] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_bad_arguments():
    with pytest.raises(ValueError):
        shard_pairs(1, 1, 0)
    with pytest.raises(ValueError):
        shard_pairs(10, 2, 2)


def _worker(rank, world, port, frames):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, n = shard_pairs(frames, world, rank)
        # every rank "interpolates" its pairs: here a marker per pair; then the same timing reduction bench.py uses
        mine = torch.zeros(frames - 1, dtype=torch.int64)
        mine[first:first + n] = rank + 1
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        owner = torch.cat([torch.full((shard_pairs(frames, world, r)[1],), r + 1) for r in range(world)])
        assert torch.equal(mine, owner), "pairs covered twice or not at all"
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 600), nprocs=2, join=True)
