"""Context2Vec: second-order (skip-gram negative-sampling) learner
(/root/reference/ADSCModel/context_embeddings.py:22-113).

Same constructor and `train(model, paths, total_nodes, alpha, node_count, chunksize)`.  The whole walk corpus is one
kernel launch (see node_embeddings.py for the workers -> mode rule).  `paths` may be an iterable of node-id
sequences (the reference's convention) or CUDA tensors straight from the device walker, in which case nothing touches
the host: `(walks, lens, G)` = CSR-row tokens of graph G, remapped on the device to table rows (Model.walks_to_rows:
CSR rows follow the first appearance of an id in the edge file, table rows follow the sorted ids), or `(walks, lens)` =
tokens that already ARE table rows.
"""
import logging as log
import time

import numpy as np

from ..utils import training_sdg_inner as K
from ..utils.embedding import paths_to_rows


class Context2Vec(object):
    def __init__(self, lr=0.1, window_size=5, workers=1, negative=5, mode=None, atomic=True, use_alias=False):
        self.lr = float(lr)
        self.workers = workers
        self.negative = negative
        self.window_size = int(window_size)
        self.mode = mode
        self.atomic = atomic
        self.use_alias = use_alias

    def _mode(self, n_tokens=0):
        return K.select_mode(self.mode, self.workers, 2 * self.window_size * int(n_tokens))

    def train(self, model, paths, total_nodes, alpha=1.0, node_count=0, chunksize=150):
        import torch
        assert model.node_embedding.dtype == torch.float32
        assert model.context_embedding.dtype == torch.float32
        log.info("O2 training model with %i workers on %i vocabulary and %i features and 'negative sampling'=%s" %
                 (self.workers, len(model.vocab), model.layer1_size, self.negative))
        if alpha <= 0.:
            return  # context_embeddings.py:58-59
        if not model.vocab:
            raise RuntimeError("you must first build vocabulary before training the model")
        if total_nodes is None:
            raise AttributeError('need the number of node')
        start = time.time()
        dev = model.node_embedding.device
        if isinstance(paths, tuple) and len(paths) in (2, 3) and hasattr(paths[0], "is_cuda"):
            walks2d, lens = paths[0], paths[1]  # device walker output: [n, L] padded with TOKEN_NONE
            if len(paths) == 3:  # CSR rows of that graph -> table rows
                walks2d = model.walks_to_rows(paths[2], walks2d)
            n_walks, L = walks2d.shape
            if model.down_sampling:  # prepare_sentences' frequent-node filter, on the device
                from .. import _lib
                _, rows_, probs_ = model.id_index()
                keep = torch.ones(model.vocab_size, dtype=torch.float32)
                keep[torch.from_numpy(rows_)] = torch.from_numpy(probs_.astype(np.float32))
                walks2d, lens = walks2d.clone(), lens.clone()
                _lib.check(_lib.load().comemb_downsample_walks(
                    _lib.ptr(walks2d), _lib.ptr(lens), n_walks, L, _lib.ptr(keep.to(dev)),
                    int(np.random.randint(0, 2 ** 31)), _lib.stream_ptr()))
            walks = walks2d.reshape(-1)
            off = torch.arange(n_walks + 1, dtype=torch.int64, device=dev) * L
        else:
            flat, off_h = paths_to_rows(model, paths)
            n_walks = off_h.size - 1
            walks = torch.from_numpy(flat.view(np.int32)).to(dev)
            off = torch.from_numpy(off_h).to(dev)
        K.check_row_tokens(walks, model.vocab_size)
        mode = self._mode(walks.numel())
        seeds = torch.from_numpy(K.draw_seeds(n_walks).view(np.int64)).to(dev)  # pyx:477, path order
        flags = 0
        alias = None
        if mode == K.MODE_HOGWILD:
            flags |= K.F_ATOMIC if self.atomic else 0
            if self.use_alias:
                alias = model.alias if getattr(model, "alias", None) is not None else model.make_alias()
        with torch.cuda.device(dev):
            tokens = K.o2_batch(model.node_embedding, model.context_embedding, walks, off, seeds, self.lr, self.negative,
                                self.window_size, model.table, alpha=alpha, mode=mode, flags=flags, alias=alias,
                                count_tokens=True,
                                max_warps=K.hogwild_concurrency(model.vocab_size, self.workers)
                                if mode == K.MODE_HOGWILD else 0)
        elapsed = time.time() - start
        log.info("training on %i nodes took %.1fs, %.0f nodes/s" % (node_count + tokens, elapsed,
                                                                     (node_count + tokens) / elapsed if elapsed else 0.0))
        return None
