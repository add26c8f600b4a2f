"""Multi-GPU sharding logic on CPU: world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from finite_difference_b200 import shard


def test_partition_uniform_and_lpt():
    parts = shard.partition(np.ones(10), 4)
    assert sorted(np.concatenate(parts).tolist()) == list(range(10))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    rng = np.random.default_rng(0)
    cost = rng.uniform(1, 100, 1001)
    parts = shard.partition(cost, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(1001))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.02
    assert shard.partition(np.ones(3), 1)[0].tolist() == [0, 1, 2]
    assert [len(p) for p in shard.partition(np.ones(0), 2)] == [0, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    r, w, _ = shard.init_distributed("gloo")
    assert (r, w) == (rank, world)
    cost = np.arange(n_total, dtype=float) % 7 + 1.0
    mine = shard.my_options(cost)
    local = {"price": mine.astype(np.float64) * 1.5, "idx": np.stack([mine, -mine], axis=1).astype(np.int32)}
    full = shard.gather_results(local, mine, n_total)
    assert np.array_equal(full["price"], np.arange(n_total) * 1.5)
    assert np.array_equal(full["idx"][:, 1], -np.arange(n_total))
    only0 = shard.gather_results(local, mine, n_total, dst=0)            # gather onto rank 0 only
    if rank == 0:
        assert np.array_equal(only0["price"], np.arange(n_total) * 1.5) and np.array_equal(only0["idx"][:, 0], np.arange(n_total))
    else:
        assert only0 is None
    # the device-side result gather (book.DeviceBookRun.packed layout), CPU tensors over gloo
    import torch
    width = shard.ShardedBookRun.shard_width(n_total, world)
    assert width % 2 == 0 and width >= len(mine)
    nf, ni = 6, 3
    packed = torch.zeros(width * (8 * nf + 4 * ni), dtype=torch.uint8)
    f = packed[: 8 * nf * width].view(torch.float64).view(nf, width)
    i = packed[8 * nf * width:].view(torch.int32).view(ni, width)
    for j in range(nf):
        f[j, : len(mine)] = torch.from_numpy(mine * 10.0 + j)
    for j in range(ni):
        i[j, : len(mine)] = torch.from_numpy((mine * 10 + j).astype(np.int32))
    for dst in (0, None):
        g = shard.PackedGather(mine, n_total, width, nf, ni, dst=dst)
        got = g(packed)
        if dst is None or rank == dst:
            F, I = got
            assert F.shape == (nf, n_total) and I.shape == (ni, n_total)
            for j in range(nf):
                assert np.array_equal(F[j], np.arange(n_total) * 10.0 + j)
            for j in range(ni):
                assert np.array_equal(I[j], np.arange(n_total) * 10 + j)
        else:
            assert got is None
    np.save(os.path.join(tmp, f"r{rank}.npy"), mine)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_gloo(tmp_path):
    n_total = 101
    mp.spawn(_worker, args=(2, _free_port(), n_total, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert sorted(np.concatenate([a, b]).tolist()) == list(range(n_total))
    assert abs(len(a) - len(b)) <= 1
