"""Host-side multi-rank logic on CPU (gloo, world_size 2): shard ranges tile the global env ids,
ownership is consistent, and the max-over-ranks timing reduction bench.py uses works."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minesweeper_ppo_b200.shard import owner_of, shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, n = shard_range(total, rank, world)
        mine = torch.zeros(total, dtype=torch.int64)
        mine[base:base + n] = rank + 1                      # claim my env ids
        dist.all_reduce(mine)                               # every id claimed exactly once
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # bench.py: max over ranks
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([n]))
        if rank == 0:
            torch.save({"claims": mine, "tmax": t, "counts": torch.cat(counts)}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [65536, 1001])
def test_shards_tile_env_ids_gloo(tmp_path, total):
    world, port, out = 2, _free_port(), str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    r = torch.load(out)
    q = total // world
    expect = torch.cat([torch.full((n,), k + 1) for k, n in enumerate(r["counts"].tolist())])
    assert torch.equal(r["claims"], expect) and int(r["counts"].sum()) == total
    assert float(r["tmax"]) == 11.0 and abs(int(r["counts"][0]) - q) <= 1


def test_shard_range_properties():
    for total, world in [(10, 4), (65536, 8), (7, 7), (4_194_304, 8), (1001, 3)]:
        seen = 0
        for rk in range(world):
            base, n = shard_range(total, rk, world)
            assert base == seen and n >= total // world
            for e in (base, base + n - 1):
                assert owner_of(e, total, world) == (rk, e - base)
            seen += n
        assert seen == total
    with pytest.raises(ValueError):
        shard_range(3, 0, 4)
