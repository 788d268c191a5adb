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
        local.append([torch.full((1, 4, r), float(i), dtype=torch.float64), torch.full((r, 5, 1), float(i) + 0.5, dtype=torch.float64)])
    allc = all_gather_cores(local, batch)  # batch < world: a rank with no problem takes dtype from the metadata
    assert len(allc) == batch
    assert all(c.dtype == torch.float64 for cores in allc for c in cores)
    for i, cores in enumerate(allc):
        r = 2 + i % 3
        assert cores[0].shape == (1, 4, r) and cores[1].shape == (r, 5, 1)
        assert float(cores[0][0, 0, 0]) == float(i) and float(cores[1][0, 0, 0]) == float(i) + 0.5
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 2, 1])
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


def _worker_generic(rank, world, port, batch):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tntorch_b200.dist import batch_sharded, shard_range

    solved = []

    def solve(p):  # stands in for ops.tt_round / ops.cp_als / cross(...).cores: ragged list of tensors per problem
        solved.append(p)
        return [torch.full((p % 3 + 1, 2), float(p)), torch.full((3, p % 2 + 1), float(-p))]

    out = batch_sharded(list(range(batch)), solve)
    lo, hi = shard_range(batch, world, rank)
    assert solved == list(range(lo, hi))  # every rank solved only its own slice, in order
    assert len(out) == batch
    for p, fac in enumerate(out):
        assert fac[0].shape == (p % 3 + 1, 2) and fac[1].shape == (3, p % 2 + 1)
        assert float(fac[0][0, 0]) == float(p) and float(fac[1][0, 0]) == float(-p)
    local_only = batch_sharded(list(range(batch)), solve, gather=False)
    assert len(local_only) == hi - lo
    dist.destroy_process_group()


def test_batch_sharded_driver_world2():
    """The driver behind round_tt_/cp_als_/cross_batch_sharded (BASELINE configs 3-5 shard over the batch)."""
    port = _free_port()
    mp.spawn(_worker_generic, args=(2, port, 7), nprocs=2, join=True)
