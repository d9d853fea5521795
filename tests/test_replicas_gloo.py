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


# ---- two ranks on ONE GPU (gloo backend on CUDA tensors) driving ReplicaTrainer with the real kernels ---------------------------
def _make_model(seed=3):
    import torch
    from comemb_b200.ADSCModel.model import Model
    from comemb_b200.utils import graph_utils as gu
    G, block = gu.sbm_graph(4000, 8, 20, p_in=0.9, seed=21)
    np.random.seed(seed)
    model = Model(G.degree(), size=128, table_size=200000, k=8)
    model.node_embedding = (model.node_embedding * 0.2).contiguous()
    return G, block, model


def _gpu_worker(rank, world, port, out, sync_every):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from comemb_b200 import replicas
    import comemb_b200.utils.training_sdg_inner as K
    K.init()
    G, block, model = _make_model()
    tr = replicas.ReplicaTrainer(model, window=5, negative=5, lr=0.05, sync_every=sync_every, flags=K.F_ATOMIC)
    for p in range(4):
        tr.step(G, 4, 30, 0.0, seed=7, pass_index=p)
    if tr.steps % sync_every:
        replicas.average_tables([model.node_embedding, model.context_embedding])
    # o1 over the sharded edge list, then o3 over sharded node rows with a planted top-1 pi
    src = np.repeat(np.arange(len(G), dtype=np.int64), np.diff(G.rowptr))
    keep = src < G.col
    edges = torch.from_numpy(np.stack([src[keep], G.col[keep].astype(np.int64)], 1).astype(np.int32)).cuda()
    o2_node, o2_ctx = model.node_embedding.clone(), model.context_embedding.clone()
    tr.sync_every = 1
    tr.step_o1(edges, seed=5)
    o1_node = model.node_embedding.clone()
    rs = np.random.RandomState(1)
    model.centroid = torch.from_numpy(rs.uniform(-0.3, 0.3, (8, 128)).astype(np.float32)).cuda()
    model.inv_covariance_mat = torch.from_numpy((rs.normal(size=(8, 128, 128)) * 0.05 + np.eye(128)).astype(np.float32)).cuda()
    comm = torch.from_numpy(np.ascontiguousarray(block, np.int32)).cuda()
    weight = torch.ones(len(G), device="cuda")
    tr.step_o3(2.0, comm=comm, weight=weight, iters=2)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"node": o2_node.cpu(), "ctx": o2_ctx.cpu(), "o1": o1_node.cpu(), "o3": model.node_embedding.cpu()},
                   os.path.join(out, "g%d.pt" % sync_every))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("sync_every", [1, 4])
def test_two_replicas_on_one_gpu_match_a_single_replica_in_quality(tmp_path, sync_every):
    """North star: "near-linear scaling ... and quality matching the reference".  Two ranks (two processes sharing cuda:0,
    gloo all-reduce on CUDA tensors) each train their half of every pass and average (every step / every 4 steps); a
    single replica trains all walks of the same passes.  Equal pair count; stated tolerance: o2 objective at most 10 % above
    the single replica's, community NMI (k-means, 5 restarts) within 0.05 when averaging every step and within 0.10 when the
    replicas meet only once after four passes (measured 0.06 there with a single-start k-means).  Then o1 over the sharded edge list stays finite and moves the
    table, and o3 over sharded node rows (disjoint shards merged by summing deltas) equals the single-process result."""
    import torch
    import torch.multiprocessing as mp
    import comemb_b200.utils.training_sdg_inner as K
    from comemb_b200 import evaluation, replicas
    K.init()
    G, block, model = _make_model()
    tr = replicas.ReplicaTrainer(model, window=5, negative=5, lr=0.05, sync_every=1, flags=K.F_ATOMIC)
    for p in range(4):
        tr.step(G, 4, 30, 0.0, seed=7, pass_index=p)
    from comemb_b200.utils import graph_utils as gu
    ew, el = gu.build_deepwalk_corpus(G, 1, 30, alpha=0.0, seed=99, mode=gu.MODE_HOGWILD, return_device=True)
    eoff = torch.arange(ew.shape[0] + 1, dtype=torch.int64, device="cuda") * 30
    l1, n1 = K.o2_pos_loss(model.node_embedding, model.context_embedding, ew.reshape(-1), eoff, 5)
    q1 = evaluation.community_nmi(model.node_embedding, block, k=8, method="kmeans")
    mp.spawn(_gpu_worker, args=(2, _free_port(), str(tmp_path), sync_every), nprocs=2, join=True)
    r = torch.load(os.path.join(str(tmp_path), "g%d.pt" % sync_every))
    node2, ctx2 = r["node"].cuda(), r["ctx"].cuda()
    assert bool(torch.isfinite(node2).all()) and bool(torch.isfinite(ctx2).all())
    assert bool(torch.isfinite(r["o1"]).all()) and not torch.equal(r["o1"], r["node"])  # the sharded o1 epoch moved the table
    G2, block2, m2 = _make_model()
    l2, n2 = K.o2_pos_loss(node2, ctx2, ew.reshape(-1), eoff, 5)
    q2 = evaluation.community_nmi(node2, block, k=8, method="kmeans")
    assert n1 == n2 and q1 > 0.8, (q1, q2)
    assert q2 >= q1 - (0.05 if sync_every == 1 else 0.10), (q1, q2)
    # not worse than the single replica by more than 10 % (averaged replicas often end LOWER: measured 1.06 vs 1.19)
    assert 0.5 * (l1 / n1) <= l2 / n2 <= 1.10 * (l1 / n1), (l1 / n1, l2 / n2)
    # o3 over node shards == the same step in one process
    rs = np.random.RandomState(1)
    m2.node_embedding = r["o1"].cuda()
    m2.centroid = torch.from_numpy(rs.uniform(-0.3, 0.3, (8, 128)).astype(np.float32)).cuda()
    m2.inv_covariance_mat = torch.from_numpy((rs.normal(size=(8, 128, 128)) * 0.05 + np.eye(128)).astype(np.float32)).cuda()
    comm = torch.from_numpy(np.ascontiguousarray(block, np.int32)).cuda()
    replicas.ReplicaTrainer(m2, 5, 5, 0.05).step_o3(2.0, comm=comm, weight=torch.ones(len(G), device="cuda"), iters=2)
    assert torch.allclose(m2.node_embedding.cpu(), r["o3"], rtol=1e-5, atol=1e-6)
    assert not torch.equal(r["o3"], r["o1"])
