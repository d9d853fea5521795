"""Row-partitioned embedding tables across the GPUs of one node (partitioning B).

Each rank (one process per GPU) owns a contiguous block of `rows_per_shard` rows of the node and the context table.
The shards of the peers are mapped into this process through CUDA IPC (torch's storage sharing) and peer access is
enabled, so the Hogwild o2 kernel (csrc/sgns_hogwild.cu, SHARDED variant) gathers remote rows and scatters
`red.global.add.v4.f32` updates straight over NVLink / NVSwitch: the exchange is fused into the SGD kernel, no
collective runs on the data path.  Every rank trains on its own shard of the walk stream; all ranks update the one
distributed table concurrently (distributed Hogwild).  The negative table and the walks are rank-local.
"""
import ctypes

import numpy as np

from . import _lib
from .utils import training_sdg_inner as K


def rows_per_shard(n_rows, world):
    return (int(n_rows) + int(world) - 1) // int(world)


class ShardedTables(object):
    def __init__(self, n_rows, size, group=None, local_only=False, n_local_shards=None):
        """local_only=True (tests): all `n_local_shards` shards live on this device, no process group needed."""
        import torch
        import torch.distributed as dist
        self.n_rows, self.size = int(n_rows), int(size)
        if size != 128:
            raise K.ComembError("row-partitioned o2 is built for size 128")
        self.local_only = local_only
        if local_only:
            self.rank, self.world = 0, int(n_local_shards or 2)
        else:
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise K.ComembError("at most 8 shards (one node)")
        self.rps = rows_per_shard(n_rows, self.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.node_shards, self.ctx_shards = [None] * self.world, [None] * self.world
        if local_only:
            for s in range(self.world):
                self.node_shards[s] = torch.zeros((self.rps, size), dtype=torch.float32, device=dev)
                self.ctx_shards[s] = torch.zeros((self.rps, size), dtype=torch.float32, device=dev)
        else:
            mine_n = torch.zeros((self.rps, size), dtype=torch.float32, device=dev)
            mine_c = torch.zeros((self.rps, size), dtype=torch.float32, device=dev)
            self.node_shards[self.rank], self.ctx_shards[self.rank] = mine_n, mine_c
            # export: torch's storage sharing yields (device, 64-byte cudaIpcMemHandle of the containing allocation,
            # size, byte offset of this storage inside it, ...); the peers map it into THEIR device context
            exp = []
            for t in (mine_n, mine_c):
                info = t.untyped_storage()._share_cuda_()
                raw = bytes(info[1])
                # newer torch prefixes the 64-byte cudaIpcMemHandle_t with a version byte and an allocation-type
                # byte ('c' = plain cudaMalloc segment, the only kind that can be opened with cudaIpcOpenMemHandle)
                if len(raw) > 64:
                    if raw[-65:-64] not in (b"c", b""):
                        raise K.ComembError("shard memory is not a cudaMalloc segment (expandable segments?)")
                    raw = raw[-64:]
                exp.append((raw, int(info[3]) + t.storage_offset() * 4))
            handles = [None] * self.world
            dist.all_gather_object(handles, (dev.index, exp), group=group)
            self._peer_ptrs = {}
            opened = {}  # one cudaIpcOpenMemHandle per exported allocation (both tables may share a segment)
            for s, (peer_dev, (hn, hc)) in enumerate(handles):
                if s == self.rank:
                    continue
                ptrs = []
                for handle, offset in (hn, hc):
                    if (s, handle) not in opened:
                        out = ctypes.c_void_p()
                        _lib.check(_lib.load().comemb_ipc_open(handle, 0, ctypes.byref(out)))
                        opened[(s, handle)] = out.value
                    ptrs.append(opened[(s, handle)] + offset)
                self._peer_ptrs[s] = tuple(ptrs)
            dist.barrier(group)
        arr = ctypes.c_void_p * 8
        node_p = [self.node_shards[s].data_ptr() if self.node_shards[s] is not None else self._peer_ptrs[s][0]
                  for s in range(self.world)]
        ctx_p = [self.ctx_shards[s].data_ptr() if self.ctx_shards[s] is not None else self._peer_ptrs[s][1]
                 for s in range(self.world)]
        self._node_ptrs = arr(*node_p + [None] * (8 - self.world))
        self._ctx_ptrs = arr(*ctx_p + [None] * (8 - self.world))
        self.group = group

    @property
    def local_node(self):
        return self.node_shards[self.rank]

    @property
    def local_ctx(self):
        return self.ctx_shards[self.rank]

    def load_rows(self, node_full=None, ctx_full=None):
        """Fill the shards this process owns from full host/device tables (numpy or torch [n_rows, size])."""
        import torch
        owned = range(self.world) if self.local_only else [self.rank]
        for s in owned:
            lo, hi = s * self.rps, min(self.n_rows, (s + 1) * self.rps)
            for full, shard in ((node_full, self.node_shards[s]), (ctx_full, self.ctx_shards[s])):
                if full is not None and hi > lo:
                    src = torch.as_tensor(full[lo:hi]) if not hasattr(full, "is_cuda") else full[lo:hi]
                    shard[:hi - lo].copy_(src)

    def local_negative_table(self, counts, table_size, power=0.75):
        """Unigram^power table over THIS rank's rows only (values are global row ids): with shard-local negatives
        only the node row and the positive context row of a pair can be remote, which cuts the NVLink bytes per pair
        by about 5/7.  Changes the sampling distribution (each rank samples its own rows), so it is a
        Hogwild-mode option for graphs that do not fit one GPU, not a parity mode.  `counts`: degrees of all rows."""
        import torch
        lo, hi = self.rank * self.rps, min(self.n_rows, (self.rank + 1) * self.rps)
        w = np.asarray(counts[lo:hi], np.float64) ** power
        cum = np.cumsum(w) / w.sum()
        slots = (np.arange(table_size, dtype=np.float64) + 0.5) / table_size
        rows = lo + np.searchsorted(cum, slots).clip(0, hi - lo - 1)
        return torch.from_numpy(rows.astype(np.int32)).to(self.local_node.device)

    def gather(self):
        """(node, ctx) full tables on this device (for evaluation / tests)."""
        import torch
        import torch.distributed as dist
        if self.local_only:
            return (torch.cat(self.node_shards)[: self.n_rows], torch.cat(self.ctx_shards)[: self.n_rows])
        out = []
        for mine in (self.local_node, self.local_ctx):
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            out.append(torch.cat(parts)[: self.n_rows])
        return out[0], out[1]

    def o2(self, walks, walk_off, seeds, lr, negative, window, table, alpha=1.0, base_seed=0, count_tokens=False):
        """Hogwild o2 over this rank's walks against the distributed tables (red.add scatter over NVLink)."""
        import torch
        _lib.ensure_init()
        flags = 0 if seeds is not None else K.F_SEED_HASH
        tok = torch.zeros(1, dtype=torch.int64, device=walks.device) if count_tokens else None
        st = _lib.load().comemb_o2_walks_sharded(
            ctypes.cast(self._node_ptrs, ctypes.c_void_p), ctypes.cast(self._ctx_ptrs, ctypes.c_void_p), self.world,
            self.rps, self.size, _lib.ptr(walks), _lib.ptr(walk_off), int(walk_off.numel()) - 1, _lib.ptr(seeds),
            int(base_seed), _lib.ptr(table), table.numel(), int(window), int(negative), float(lr), float(alpha),
            int(flags), _lib.ptr(tok), _lib.stream_ptr())
        _lib.check(st)
        return int(tok.item()) if count_tokens else None
