"""Row-partitioned o2 across 2+ GPUs of one node (run under torchrun on a multi-GPU box):
   torchrun --nproc-per-node 2 scripts/sharded_p2p_check.py
Checks (1) rank 0 alone, walking rows of every shard, reproduces the flat single-GPU result bit for bit (remote rows
read and red.add-updated over NVLink); (2) all ranks together: finite, moving, token counts right; prints throughput."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import comemb_b200.utils.training_sdg_inner as K
    from comemb_b200.sharded import ShardedTables
    K.init()
    n, d, L, W, neg = 100000, 128, 80, 10, 5
    rs = np.random.RandomState(0)
    node = (rs.uniform(-1, 1, (n, d)) * 0.05).astype(np.float32)
    ctx = (rs.uniform(-1, 1, (n, d)) * 0.05).astype(np.float32)
    table = torch.from_numpy(np.sort(rs.randint(1, n, 2000000)).astype(np.int32)).cuda()
    st = ShardedTables(n, d)
    st.load_rows(node, ctx)
    dist.barrier()
    # (1) exactness over NVLink: rank 0 runs 6 walks one at a time; everyone else idles
    walks = torch.from_numpy(rs.randint(0, n, (6, L)).astype(np.int32)).cuda()
    off1 = torch.tensor([0, L], dtype=torch.int64, device="cuda")
    if rank == 0:
        fn, fc = torch.from_numpy(node).cuda(), torch.from_numpy(ctx).cuda()
        for i in range(6):
            sd = torch.tensor([1234567 + i], dtype=torch.int64, device="cuda")
            K.o2_batch(fn, fc, walks[i], off1, sd, 0.025, neg, W, table, mode=K.MODE_HOGWILD, flags=K.F_ATOMIC)
            st.o2(walks[i], off1, sd, 0.025, neg, W, table)
        torch.cuda.synchronize()
    dist.barrier()
    gn, gc = st.gather()
    if rank == 0:
        ok = torch.equal(gn, fn) and torch.equal(gc, fc)
        print("rank0 sharded-over-NVLink == flat single GPU:", ok, flush=True)
        assert ok
    dist.barrier()
    # (2) all ranks train concurrently on their own walks; throughput
    nw = 50000
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    mywalks = torch.randint(0, n, (nw, L), device="cuda", generator=g, dtype=torch.int32)
    off = torch.arange(nw + 1, dtype=torch.int64, device="cuda") * L
    for rep in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tok = st.o2(mywalks.reshape(-1), off, None, 0.025, neg, W, table, base_seed=7 + rep, count_tokens=True)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            print("row-partitioned o2, %d GPUs: %.3g pair-updates/s total (%.1f ms)" % (world, world * nw * 1490 / dt,
                                                                                      dt * 1e3), flush=True)
        assert tok == nw * L
    gn, gc = st.gather()
    assert torch.isfinite(gn).all() and torch.isfinite(gc).all()
    if rank == 0:
        print("moved rows:", int((gn.cpu() != torch.from_numpy(node)).any(1).sum()), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
