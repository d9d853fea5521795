"""Multi-GPU data parallelism of the SGD path: replicated tables, sharded walk/edge stream, periodic averaging.

One process per GPU (torch.distributed, backend nccl over NVLink/NVSwitch).  Every rank owns full copies of the node
and context tables and trains on its own contiguous shard of the walk space; every `sync_every` steps the replicas are
averaged:  w <- (1/G) * sum_g w_g   (== snapshot + mean of the per-rank deltas).  The sum is an NCCL all-reduce
(in-switch NVLS reduction on NVSwitch systems); the 1/G scaling is csrc's comemb_scale kernel on CUDA tensors.
The reference has no distributed path: this is new, Hogwild-mode-only functionality.
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
    works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tables]
    for w, t in zip(works, tables):
        w.wait()
        if t.is_cuda:
            with torch.cuda.device(t.device):
                _lib.check(_lib.load().comemb_scale(t.data_ptr(), t.numel(), 1.0 / world, _lib.stream_ptr()))
        else:  # host tensors (gloo): used by the CPU tests of this module's logic
            t.mul_(1.0 / world)


class ReplicaTrainer(object):
    """Hogwild o2 epochs on one rank's shard with periodic averaging."""

    def __init__(self, model, window, negative, lr, sync_every=1, flags=0, group=None):
        import torch.distributed as dist
        self.model, self.window, self.negative, self.lr = model, window, negative, lr
        self.sync_every, self.flags, self.group = sync_every, flags, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.steps = 0

    def step(self, G, num_paths, path_length, alpha_restart, seed, pass_index):
        """One pass: this rank walks its shard of pass `pass_index` on the device and trains on it."""
        import torch
        from .utils import graph_utils as gu
        from .utils import training_sdg_inner as K
        n = len(G)
        first, count = shard_range(n, self.rank, self.world)
        walks, lens = gu.build_deepwalk_corpus(G, num_paths, path_length, alpha=alpha_restart, seed=seed,
                                               mode=gu.MODE_HOGWILD, return_device=True,
                                               first_walk=pass_index * n + first, n_out=count)
        off = torch.arange(count + 1, dtype=torch.int64, device=walks.device) * path_length
        walks = self.model.walks_to_rows(G, walks)  # CSR rows -> table rows (identity for sorted dense ids)
        K.o2_batch(self.model.node_embedding, self.model.context_embedding, walks.reshape(-1), off, None, self.lr,
                   self.negative, self.window, self.model.table, mode=K.MODE_HOGWILD, flags=self.flags,
                   base_seed=seed * 1000003 + pass_index * 8191 + self.rank)
        self.steps += 1
        if self.steps % self.sync_every == 0:
            average_tables([self.model.node_embedding, self.model.context_embedding], self.group, self.world)
        return walks, lens
