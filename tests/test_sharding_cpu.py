"""CPU: the N>1 path (sharding + final gather) with world_size=2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dyobav_mpcnwta_warehouse_b200 import sharding


def test_shard_range_partitions_everything():
    for n in (1, 2, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                got.extend(range(lo, hi))
                assert (hi - lo) in (n // world, n // world + 1)
            assert got == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 4, 4)


def test_best_of_starts_picks_lowest_cost_and_ignores_nan():
    cost = torch.tensor([3.0, 1.0, 2.0, float("nan"), 5.0, 4.0])
    u = torch.arange(12, dtype=torch.float64).reshape(6, 2)
    st = torch.tensor([0, 1, 0, 3, 0, 1], dtype=torch.int32)
    c, uu, s, idx = sharding.best_of_starts(cost, u, st, starts=3)
    assert c.tolist() == [1.0, 4.0] and idx.tolist() == [1, 2]
    assert uu.tolist() == [[2.0, 3.0], [10.0, 11.0]] and s.tolist() == [1, 1]


def _worker(rank, world, port, n_scen, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_scen, rank, world)
    # every rank "solves" its slice: result row i holds the global scenario id
    local = {"u": torch.arange(lo, hi, dtype=torch.float64)[:, None].repeat(1, 4),
             "exit_status": torch.arange(lo, hi, dtype=torch.int32)}
    full = sharding.gather_results(local, n_scen)
    ok = (full["u"][:, 0].tolist() == list(map(float, range(n_scen)))
          and full["exit_status"].tolist() == list(range(n_scen)))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_scen", [8, 7])
def test_gather_results_gloo_world2(n_scen):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_scen, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
