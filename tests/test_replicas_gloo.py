"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard ranges and replica averaging."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from comemb_b200 import replicas
    base = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    node = base + rank          # replicas drifted apart by their local updates
    ctx = base * (rank + 1)
    replicas.average_tables([node, ctx])
    first, count = replicas.shard_range(101, rank, world)
    torch.save({"node": node, "ctx": ctx, "shard": (first, count)}, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_average_tables_and_shards_world2(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    base = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    for r in res:
        assert torch.allclose(r["node"], base + 0.5)      # mean of base+0, base+1
        assert torch.allclose(r["ctx"], base * 1.5)       # mean of base*1, base*2
    shards = [r["shard"] for r in res]
    assert shards[0][0] == 0 and shards[0][0] + shards[0][1] == shards[1][0] and sum(c for _, c in shards) == 101


def test_shard_range_partitions_everything():
    from comemb_b200.replicas import shard_range
    for total in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_rows_per_shard():
    from comemb_b200.sharded import rows_per_shard
    assert rows_per_shard(100000, 8) == 12500 and rows_per_shard(100001, 8) == 12501 and rows_per_shard(5, 8) == 1
    for n in (1, 34, 100000, 50000000):
        for w in (1, 2, 4, 8):
            r = rows_per_shard(n, w)
            assert r * w >= n and (r - 1) * w < n
