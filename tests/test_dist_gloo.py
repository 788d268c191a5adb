"""World-size-2 gloo test of the multi-GPU plumbing (batch sharding + final all-gather of ragged cores)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tntorch_b200.dist import all_gather_cores, shard_range

    lo, hi = shard_range(batch, world, rank)
    # problem i has data-dependent ranks (1, 2+i%3, 1): ragged shapes across ranks
    local = []
    for i in range(lo, hi):
        r = 2 + i % 3
        local.append([torch.full((1, 4, r), float(i)), torch.full((r, 5, 1), float(i) + 0.5)])
    allc = all_gather_cores(local, batch)
    assert len(allc) == batch
    for i, cores in enumerate(allc):
        r = 2 + i % 3
        assert cores[0].shape == (1, 4, r) and cores[1].shape == (r, 5, 1)
        assert float(cores[0][0, 0, 0]) == float(i) and float(cores[1][0, 0, 0]) == float(i) + 0.5
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 2])
def test_all_gather_ragged_cores_world2(batch):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, batch), nprocs=2, join=True)


def test_shard_range_partitions():
    from tntorch_b200.dist import shard_range

    for batch in (0, 1, 7, 8, 512):
        for world in (1, 2, 4, 8):
            ranges = [shard_range(batch, world, g) for g in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == batch
            assert all(ranges[g][1] == ranges[g + 1][0] for g in range(world - 1))
            assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1
