"""Node2Vec: first-order proximity learner (/root/reference/ADSCModel/node_embeddings.py:18-106).

Same constructor and `train(model, edges, chunksize, iter)` / `loss(model, edges)`.  The reference feeds one edge at a
time to train_o1 from `workers` threads; here the whole (repeated) edge list is one kernel launch:
    workers == 1 -> ORDERED mode: the reference's sequential result, bit for bit (seeds drawn from np.random in edge
                    order exactly as the single worker thread does) -- for calls of up to
                    K.ORDERED_AUTO_MAX_UPDATES pair updates; the sequential replay is latency-bound (~1e6 updates/s
                    whatever the hardware), so larger calls train lock-free unless
                    mode="ordered" is given;
    workers  > 1 -> HOGWILD mode: lock-free, one warp per edge (the reference's multi-thread semantics at GPU width).
`mode=` ("ordered" | "hogwild") overrides that choice (K.select_mode).
"""
import logging as log
import time

import numpy as np

from ..utils import training_sdg_inner as K
from ..utils.embedding import RepeatCorpusNTimes, paths_to_rows


class Node2Vec(object):
    def __init__(self, lr=0.2, workers=1, negative=0, mode=None, atomic=True):
        self.workers = workers
        self.lr = float(lr)
        self.negative = negative
        self.window_size = 1
        self.mode = mode
        self.atomic = atomic

    def _mode(self, n_edges=0):
        return K.select_mode(self.mode, self.workers, 2 * int(n_edges))

    def loss(self, model, edges):
        """-sum log sigmoid(x_u . x_v) over in-vocabulary edges (node_embeddings.py:26-31), on the device."""
        import torch
        flat, off = paths_to_rows(model, edges)
        assert (np.diff(off) == 2).all(), "edges have to be done by 2 nodes"
        e = torch.from_numpy(flat.astype(np.int64).reshape(-1, 2)).to(model.node_embedding.device)
        x = model.node_embedding
        z = (x[e[:, 0]].double() * x[e[:, 1]].double()).sum(1)
        return float(-torch.nn.functional.logsigmoid(z).sum().item())

    def train(self, model, edges, chunksize=150, iter=1):
        import torch
        assert model.node_embedding.dtype == torch.float32
        if not model.vocab:
            raise RuntimeError("you must first build vocabulary before training the model")
        log.info("O1 training model with %i workers on %i vocabulary and %i features and 'negative sampling'=%s" %
                 (self.workers, len(model.vocab), model.layer1_size, self.negative))
        start = time.time()
        flat, off = paths_to_rows(model, RepeatCorpusNTimes(edges, iter))
        lens = np.diff(off)
        if not (lens == 2).all():
            # an edge shortened by OOV filtering makes the reference read uninitialised indexes
            keep = np.repeat(lens == 2, lens)
            flat = flat[keep]
        n_edges = flat.size // 2
        mode = self._mode(n_edges)
        dev = model.node_embedding.device
        e = torch.from_numpy(flat.view(np.int32)).to(dev)
        K.check_row_tokens(e, model.vocab_size, "edge endpoints")
        flags = 0
        if mode == K.MODE_ORDERED:
            seeds = torch.from_numpy(K.draw_seeds(n_edges).view(np.int64)).to(dev)  # pyx:427, edge order
            stride = 0
        else:
            seeds = torch.from_numpy(K.draw_seeds(n_edges).view(np.int64)).to(dev)
            flags |= K.F_ATOMIC if self.atomic else 0
            stride = _coprime_stride(n_edges)
        with torch.cuda.device(dev):
            K.o1_batch(model.node_embedding, e, seeds, self.lr, self.negative, model.table, mode=mode, flags=flags,
                       edge_stride=stride,
                       max_warps=K.hogwild_concurrency(model.vocab_size, self.workers) if mode == K.MODE_HOGWILD else 0)
            torch.cuda.current_stream().synchronize()
        elapsed = time.time() - start
        log.info("training on %i words took %.1fs, %.0f words/s" % (2 * n_edges, elapsed,
                                                                     2 * n_edges / elapsed if elapsed else 0.0))


def _coprime_stride(n):
    """A multiplier near n*0.618 coprime to n: consecutive warps then work on far-apart edges."""
    from math import gcd
    if n < 3:
        return 0
    s = max(2, int(n * 0.6180339887))
    while gcd(s, n) != 1:
        s += 1
    return s
