"""Multi-GPU data parallelism of the SGD path: replicated tables, sharded walk / edge / node streams, periodic averaging.

One process per GPU (torch.distributed, backend nccl over NVLink/NVSwitch).  Every rank owns full copies of the node
and context tables and trains on its own contiguous shard of the work (SURVEY 8e partition A):
    o2 / fused pass : the walk space of a pass (the walker generates only this rank's walks),
    o1              : the edge list (node_embeddings.py:70-71),
    o3              : the node rows (community_embeddings.py:61-77: a row's update depends on that row only, so the node
                      shards are disjoint and the "average" of G replicas of which one moved a row is handled exactly by
                      summing the deltas);
every `sync_every` steps the replicas are averaged:  w <- (1/G) * sum_g w_g   (== snapshot + mean of the per-rank deltas).
With NCCL the reduction is a single ncclAvg all-reduce per table (in-switch NVLS reduction on NVSwitch systems, the 1/G
folded into the collective); other backends (gloo: the CPU tests and the one-GPU two-process test) sum and scale with
csrc's comemb_scale.  The reference has no distributed path: this is new, Hogwild-mode-only.
"""
from . import _lib


def shard_range(total, rank, world):
    """Contiguous shard [first, first+count) of `total` work items for `rank` of `world`; sizes differ by at most 1."""
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def average_tables(tables, group=None, world=None):
    """In-place replica average of a list of tensors (all ranks must call with same-shaped tensors)."""
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return
    if dist.get_backend(group) == "nccl":  # the 1/G scale rides inside the collective
        works = [dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=True) for t in tables]
        for w in works:
            w.wait()
        return
    works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tables]
    for w, t in zip(works, tables):
        w.wait()
        if t.is_cuda:
            with torch.cuda.device(t.device):
                _lib.check(_lib.load().comemb_scale(t.data_ptr(), t.numel(), 1.0 / world, _lib.stream_ptr()))
        else:  # host tensors (gloo): used by the CPU tests of this module's logic
            t.mul_(1.0 / world)


def sum_deltas(table, snapshot, group=None, world=None):
    """table <- snapshot + sum_g (table_g - snapshot): for steps whose ranks update DISJOINT rows (o3 over node shards),
    where averaging would shrink every update by 1/G."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return
    table.sub_(snapshot)
    dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    table.add_(snapshot)


class ReplicaTrainer(object):
    """Hogwild epochs of o1 / o2 / o3 / the fused pass on one rank's shard with periodic averaging."""

    def __init__(self, model, window, negative, lr, sync_every=1, flags=0, group=None):
        import torch.distributed as dist
        self.model, self.window, self.negative, self.lr = model, window, negative, lr
        self.sync_every, self.flags, self.group = sync_every, flags, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.steps = 0

    def _maybe_sync(self, tables):
        self.steps += 1
        if self.steps % self.sync_every == 0:
            average_tables(tables, self.group, self.world)

    def _walks(self, G, num_paths, path_length, alpha_restart, seed, pass_index):
        import torch
        from .utils import graph_utils as gu
        n = len(G)
        first, count = shard_range(n, self.rank, self.world)
        walks, lens = gu.build_deepwalk_corpus(G, num_paths, path_length, alpha=alpha_restart, seed=seed,
                                               mode=gu.MODE_HOGWILD, return_device=True,
                                               first_walk=pass_index * n + first, n_out=count)
        off = torch.arange(count + 1, dtype=torch.int64, device=walks.device) * path_length
        walks = self.model.walks_to_rows(G, walks)  # CSR rows -> table rows (identity for sorted dense ids)
        return walks, lens, off

    def step(self, G, num_paths, path_length, alpha_restart, seed, pass_index):
        """One o2 pass: this rank walks its shard of pass `pass_index` on the device and trains on it."""
        from .utils import training_sdg_inner as K
        walks, lens, off = self._walks(G, num_paths, path_length, alpha_restart, seed, pass_index)
        K.o2_batch(self.model.node_embedding, self.model.context_embedding, walks.reshape(-1), off, None, self.lr,
                   self.negative, self.window, self.model.table, mode=K.MODE_HOGWILD, flags=self.flags,
                   base_seed=seed * 1000003 + pass_index * 8191 + self.rank)
        self._maybe_sync([self.model.node_embedding, self.model.context_embedding])
        return walks, lens

    def step_fused(self, G, num_paths, path_length, alpha_restart, seed, pass_index, lambda1, lambda2, comm=None,
                   weight=None):
        """One pass of the legacy fused pass (o3 gradient + SGNS per pair) on this rank's walk shard; pi from the model
        (dense) or in top-1 form (comm, weight)."""
        from .utils import training_sdg_inner as K
        m = self.model
        walks, lens, off = self._walks(G, num_paths, path_length, alpha_restart, seed, pass_index)
        base = seed * 1000003 + pass_index * 8191 + self.rank
        if comm is not None:
            K.sg_batch_top1(m.node_embedding, m.context_embedding, walks.reshape(-1), off, None, None, self.lr,
                            self.negative, self.window, m.table, m.centroid, m.inv_covariance_mat, comm, weight, lambda1,
                            lambda2, flags=self.flags, base_seed=base)
        else:
            K.sg_batch(m.node_embedding, m.context_embedding, walks.reshape(-1), off, None, None, self.lr, self.negative,
                       self.window, m.table, m.centroid, m.inv_covariance_mat, m.pi, lambda1, lambda2, 0,
                       mode=K.MODE_HOGWILD, flags=self.flags, base_seed=base)
        self._maybe_sync([m.node_embedding, m.context_embedding])
        return walks, lens

    def step_o1(self, edges, seed, pass_index=0):
        """One o1 epoch over this rank's shard of the edge list (uint32 CUDA tensor [E, 2] of table rows)."""
        from .utils import training_sdg_inner as K
        from .ADSCModel.node_embeddings import _coprime_stride
        first, count = shard_range(edges.shape[0], self.rank, self.world)
        mine = edges[first:first + count].contiguous()
        K.o1_batch(self.model.node_embedding, mine, None, self.lr, self.negative, self.model.table, mode=K.MODE_HOGWILD,
                   flags=self.flags, base_seed=seed * 1000003 + pass_index * 8191 + self.rank,
                   edge_stride=_coprime_stride(count))
        self._maybe_sync([self.model.node_embedding])

    def step_o3(self, beta, comm=None, weight=None, iters=1):
        """Community2Vec.train over this rank's shard of the node rows; the shards are disjoint, so the replicas are
        merged by summing the per-rank deltas (exactly the single-process result)."""
        import torch
        from .utils import training_sdg_inner as K
        m = self.model
        n = m.node_embedding.shape[0]
        first, count = shard_range(n, self.rank, self.world)
        rows = torch.arange(first, first + count, dtype=torch.int32, device=m.node_embedding.device)
        snap = m.node_embedding.clone() if self.world > 1 else None
        inv_t = K.transpose_blocks(m.inv_covariance_mat.contiguous())
        if comm is None:
            K.o3_batch(m.node_embedding, rows, m.centroid, inv_t, m.pi.contiguous(), beta, self.lr, iters=iters)
        else:
            K.o3_batch_top1(m.node_embedding, rows, m.centroid, inv_t, comm, weight, beta, self.lr, iters=iters)
        if self.world > 1:
            sum_deltas(m.node_embedding, snap, self.group, self.world)
