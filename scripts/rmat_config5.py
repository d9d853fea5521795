"""BASELINE configs[4] shape: R-MAT graph (a,b,c,d = 0.57,0.19,0.19,0.05), d=128, row-partitioned embedding tables
across the GPUs of one node, Hogwild o2 with shard-local negatives; everything (graph generation, CSR, walks, SGD) on
the devices.  Run under torchrun:

    torchrun --nproc-per-node 8 scripts/rmat_config5.py --nodes 50000000 --edges 1000000000

Every rank builds the same CSR (same seed) on its own GPU; tables are split into contiguous row blocks
(sharded.ShardedTables) and each rank trains on its own walks, reading/updating remote rows over NVLink inside the
kernel.  Prints pair-updates/s (device-timed, max over ranks) and the memory footprint.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rmat_csr(n, n_edges, seed, chunk=100_000_000):
    """Undirected R-MAT multigraph as CSR on the current device (row = node, neighbours unsorted within a row)."""
    scale = int(np.ceil(np.log2(n)))
    g = torch.Generator(device="cuda").manual_seed(seed)
    a, b, c = 0.57, 0.19, 0.19
    us, vs = [], []
    done = 0
    while done < n_edges:
        m = int(min(chunk, n_edges - done))
        u = torch.zeros(m, dtype=torch.int64, device="cuda")
        v = torch.zeros(m, dtype=torch.int64, device="cuda")
        for _ in range(scale):
            r = torch.rand(m, device="cuda", generator=g)
            ub = r >= (a + b)
            vb = ((r >= a) & (r < a + b)) | (r >= a + b + c)
            u = (u << 1) | ub
            v = (v << 1) | vb
        us.append((u % n).to(torch.int32))
        vs.append((v % n).to(torch.int32))
        done += m
        del u, v, r, ub, vb
    u, v = torch.cat(us), torch.cat(vs)
    del us, vs
    src = torch.cat([u, v])
    dst = torch.cat([v, u])
    del u, v
    deg = torch.bincount(src, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(deg, 0, out=rowptr[1:])
    order = torch.argsort(src)
    del src
    col = dst[order].contiguous()
    del dst, order
    torch.cuda.empty_cache()
    return rowptr, col, deg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=50_000_000)
    ap.add_argument("--edges", type=int, default=1_000_000_000)
    ap.add_argument("--walks-per-step", type=int, default=200_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import comemb_b200.utils.training_sdg_inner as K
    from comemb_b200 import _lib
    from comemb_b200.sharded import ShardedTables
    K.init()
    n, d, L, W, neg = args.nodes, 128, 80, 10, 5
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    rowptr, col, deg = rmat_csr(n, args.edges, seed=12345)
    t1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("R-MAT CSR on device: %d nodes, %d adjacency entries, %.1f s, %d isolated nodes, max degree %d" % (
            n, col.numel(), t0.elapsed_time(t1) / 1e3, int((deg == 0).sum()), int(deg.max())), flush=True)
    st = ShardedTables(n, d)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    st.local_node.copy_((torch.rand(st.local_node.shape, device="cuda", generator=g) - 0.5) * 0.1)
    counts = deg.cpu().numpy().astype(np.float64)
    table = st.local_negative_table(np.maximum(counts, 1e-3), 5_000_000)
    col32 = col.to(torch.int32)
    del col
    dist.barrier()
    nws = args.walks_per_step
    walks = torch.empty((nws, L), dtype=torch.int32, device="cuda")
    lens = torch.empty(nws, dtype=torch.int32, device="cuda")
    off = torch.arange(nws + 1, dtype=torch.int64, device="cuda") * L
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    i = np.arange(L + 1)[:, None]
    pos = np.arange(L)[None, :]
    pairs_of_len = np.where(pos < i, np.minimum(i, pos + W + 1) - np.maximum(0, pos - W) - 1, 0).sum(1)
    pairs_lut = torch.from_numpy(pairs_of_len.astype(np.int64)).cuda()
    times, pairs = [], 0
    for s in range(args.warmup + args.steps):
        first = ((s * world + rank) * nws) % max(1, n - nws)
        _lib.check(lib.comemb_walks_csr(rowptr.data_ptr(), col32.data_ptr(), n, 1 << 10, L, 0.0, 777, K.MODE_HOGWILD,
                                        first, nws, walks.data_ptr(), lens.data_ptr(), stream.cuda_stream))
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.o2(walks.reshape(-1), off, None, 0.025, neg, W, table, base_seed=1000 * s + rank)
        e1.record()
        torch.cuda.synchronize()
        if s >= args.warmup:
            times.append(e0.elapsed_time(e1))
            pairs += int(pairs_lut[lens.long()].sum().item())
    t = torch.tensor([sum(times), float(pairs)], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    finite = bool(torch.isfinite(st.local_node).all()) and bool(torch.isfinite(st.local_ctx).all())
    if rank == 0:
        print(json.dumps({"config": "R-MAT %d nodes / %d edges, d=128, row-partitioned over %d GPUs, shard-local negatives"
                          % (n, args.edges, world), "pair_updates_per_sec": float(tsum[1]) / (float(tmax[0]) * 1e-3),
                          "pairs_per_step_all_ranks": float(tsum[1]) / args.steps, "ms_per_step": float(tmax[0]) / args.steps,
                          "table_bytes_per_gpu": 2 * st.rps * d * 4, "csr_bytes_per_gpu": col32.numel() * 4 + rowptr.numel() * 8,
                          "max_memory_allocated_gb": torch.cuda.max_memory_allocated() / 1e9, "finite": finite}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
