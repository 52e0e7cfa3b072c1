"""CPU, world_size 2, gloo: the N > 1 host logic of the batched path (sharding of
independent frames, cross-rank agreement check, max-over-ranks clock)."""
import os
import socket

import numpy as np
import pytest

from libvips_b200 import shard


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank thumbnails its own block of an 11-frame batch with the ORACLE standing in for
        # the kernel (no GPU here): the union must equal the single-process result
        from oracle import pyoracle
        frames = np.random.default_rng(1234).integers(0, 256, (11, 64, 96, 4), dtype=np.uint8)
        lo, hi = shard.shard_range(len(frames), rank, world)
        mine = np.stack([pyoracle.thumbnail_image(f, 16) for f in frames[lo:hi]])
        csum = torch.tensor([int(mine.astype(np.int64).sum())])
        total = csum.clone()
        dist.all_reduce(total)
        # the shared frame (index 0 everywhere) must agree bit for bit across ranks
        shared = torch.from_numpy(pyoracle.thumbnail_image(frames[0], 16).astype(np.int64)).sum().reshape(1)
        agree = shard.all_agree(dist, shared)
        disagree = shard.all_agree(dist, torch.tensor([rank]))
        t = shard.max_over_ranks(dist, torch.tensor([1.0 + rank], dtype=torch.float64))
        results[rank] = (lo, hi, int(total), agree, disagree, float(t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    from oracle import pyoracle
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    frames = np.random.default_rng(1234).integers(0, 256, (11, 64, 96, 4), dtype=np.uint8)
    want = int(np.stack([pyoracle.thumbnail_image(f, 16) for f in frames]).astype(np.int64).sum())
    assert results[0][:2] == (0, 6) and results[1][:2] == (6, 11)
    for r in range(world):
        assert results[r][2] == want       # union of the shards == whole batch
        assert results[r][3] is True       # shared frame agrees
        assert results[r][4] is False      # and the check does notice disagreement
        assert results[r][5] == 2.0        # slowest rank's clock
    assert shard.aggregate_rate(10, 2, 2.0) == 10.0
