"""CPU, world_size 2 over gloo: the N>1 host logic of the path (batch sharding without a data-path collective,
max-over-ranks timing, ragged gather of decoded joints)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hrnet_b200.parallel import GradAllReduce, gather_joints, max_over_ranks, shard_range


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 8, 64, 129):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        g = torch.Generator().manual_seed(5)
        all_joints = torch.rand(n_total, 21, 2, generator=g)          # same on every rank
        local = all_joints[lo:hi].clone()                            # "decoded on this rank's shard"
        out = gather_joints(local, n_total)
        assert torch.equal(out, all_joints)
        ms = max_over_ranks(10.0 + 5.0 * rank)
        assert ms == 10.0 + 5.0 * (world - 1)
        # weak-scaling aggregate the bench reports: units of all ranks / slowest rank's time
        value = n_total / (ms / 1e3)
        assert abs(value - n_total / 0.015) < 1e-6
        # training exchange step: bucketed sum-all-reduce of the flat gradient buffer; mean via the optimizer's scale
        for nb in (1, 3):
            grads = torch.arange(1000, dtype=torch.float32) * (rank + 1)
            ar = GradAllReduce(grads.numel(), n_buckets=nb)
            assert ar.bounds[0][0] == 0 and ar.bounds[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(ar.bounds, ar.bounds[1:]))
            ar(grads)
            assert torch.equal(grads, torch.arange(1000, dtype=torch.float32) * sum(range(1, world + 1)))
            assert ar.mean_scale == 1.0 / world
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_and_timing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 13), nprocs=2, join=True)
