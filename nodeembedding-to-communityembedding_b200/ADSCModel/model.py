"""Model state with the reference's attribute names (/root/reference/ADSCModel/model.py), tables on the GPU.

    vocab, vocab_size, layer1_size, node_embedding [N,d], context_embedding [N,d], centroid [K,d],
    covariance_mat / inv_covariance_mat [K,d,d], pi [N,K], table [table_size] uint32, k, ground_true

`node_embedding` etc. are torch CUDA tensors (float32); `table` is a CUDA int32 tensor holding the uint32 values
of model.py:97-122 (bit-exact, built by comemb_make_table).  Initial values come from the same legacy `np.random`
draws as the reference (model.py:86), so a seeded run starts from identical tables.
"""
import logging as log
import pickle
from os import makedirs
from os.path import exists, join as path_join

import numpy as np

from .. import _lib
from ..utils.embedding import Vocab
from ..utils.IO_utils import load_ground_true


class Model(object):
    def __init__(self, nodes_degree, size=2, down_sampling=0, seed=1, table_size=100000000, path_labels="data/",
                 input_file=None, k=None, device="cuda"):
        self.down_sampling = down_sampling
        self.seed = seed
        self.table_size = int(table_size)
        if size % 4 != 0:
            log.warning("consider setting layer size to a multiple of 4 for greater performance")
        self.layer1_size = int(size)
        self.device = device
        if nodes_degree is None:
            raise Exception("Model not initialized, need the nodes degree")
        self.build_vocab_(dict(nodes_degree))
        if input_file is not None:
            self.ground_true, self.k = load_ground_true(path=path_labels, file_name=input_file)
        else:
            self.ground_true, self.k = None, int(k or 1)
        if k is not None:
            self.k = int(k)
        self.reset_weights()
        self.make_table()

    # ---- vocabulary (model.py:52-81) -------------------------------------------------------------------------------------
    def build_vocab_(self, vocab):
        self.vocab = {}
        for node_idx, (node, count) in enumerate(sorted(vocab.items(), key=lambda itm: itm[0])):
            v = Vocab()
            v.count = count
            v.index = node_idx
            self.vocab[node] = v
        assert min(self.vocab.keys()) == 1  # model.py:66 (make_table's id-as-row convention depends on it)
        self.precalc_sampling()
        self._id_index = None

    def precalc_sampling(self):
        if self.down_sampling:
            total_nodes = sum(v.count for v in self.vocab.values())
            threshold_count = float(self.down_sampling) * total_nodes
        for v in self.vocab.values():
            prob = (np.sqrt(v.count / threshold_count) + 1) * (threshold_count / v.count) if self.down_sampling else 1.0
            v.sample_probability = min(prob, 1.0)

    def id_index(self):
        """(sorted node ids, their rows, their sample probabilities) as arrays, for vectorised id -> row mapping."""
        if self._id_index is None:
            ids = np.fromiter(self.vocab.keys(), dtype=np.int64, count=len(self.vocab))
            order = np.argsort(ids, kind="stable")
            ids = ids[order]
            vs = list(self.vocab.values())
            rows = np.asarray([vs[i].index for i in order], np.int64)
            probs = np.asarray([vs[i].sample_probability for i in order], np.float64)
            self._id_index = (ids, rows, probs)
        return self._id_index

    def graph_row_lut(self, G):
        """Device LUT from the CSR rows of graph `G` (first-appearance order of the ids, utils/graph_utils.Graph) to the
        rows of this model's tables (rank of the node id, model.py:60-65), -1 for nodes outside the vocabulary; None when
        the two orders coincide (ids sorted and all in the vocabulary -- every synthetic generator), so that the common
        case costs nothing.  The device walker emits CSR rows; the tables are indexed by vocabulary rows."""
        import torch
        ids, rows, _ = self.id_index()
        gid = np.asarray(G.ids, np.int64)
        pos = np.searchsorted(ids, gid)
        pos[pos >= ids.size] = 0
        ok = ids[pos] == gid
        lut = np.where(ok, rows[pos], -1).astype(np.int32)
        if ok.all() and np.array_equal(lut, np.arange(gid.size, dtype=np.int32)):
            return None
        return torch.from_numpy(lut).to(self.device)

    def walks_to_rows(self, G, walks):
        """CSR-row walk tokens of `G` (TOKEN_NONE padding kept) -> table-row tokens; out-of-vocabulary nodes become
        TOKEN_NONE, which the kernels skip like the reference skips `None` (pyx:485-486)."""
        import torch
        lut = self.graph_row_lut(G)
        if lut is None:
            return walks
        ext = torch.cat([lut, torch.full((1,), -1, dtype=torch.int32, device=lut.device)])
        idx = torch.where(walks < 0, torch.full_like(walks, lut.numel()), walks).long()
        return ext[idx]

    # ---- tables (model.py:83-92) -----------------------------------------------------------------------------------------
    def reset_weights(self):
        import torch
        self.vocab_size = len(self.vocab)
        init = np.random.uniform(low=-1, high=1, size=(self.vocab_size, self.layer1_size)).astype(np.float32)
        dev = self.device
        self.node_embedding = torch.from_numpy(init).to(dev)
        self.context_embedding = torch.zeros((self.vocab_size, self.layer1_size), dtype=torch.float32, device=dev)
        self.centroid = torch.zeros((self.k, self.layer1_size), dtype=torch.float32, device=dev)
        self.covariance_mat = torch.zeros((self.k, self.layer1_size, self.layer1_size), dtype=torch.float32, device=dev)
        self.inv_covariance_mat = torch.zeros_like(self.covariance_mat)
        self.pi = torch.zeros((self.vocab_size, self.k), dtype=torch.float32, device=dev)

    # ---- negative-sampling table (model.py:97-122) ------------------------------------------------------------------------
    def make_table(self, power=0.75):
        import torch
        counts = np.asarray([self.vocab[i].count for i in sorted(self.vocab)], np.float64)
        self.table = torch.empty(self.table_size, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.table.device):
            st = _lib.load().comemb_make_table(counts.ctypes.data, counts.size, float(power), _lib.ptr(self.table),
                                               self.table_size, _lib.stream_ptr())
        _lib.check(st)
        self.alias = None

    def make_alias(self):
        """Alias table with the same distribution as `table` (Hogwild option)."""
        import torch
        self.alias = torch.empty(2 * self.vocab_size, dtype=torch.int32, device=self.device)
        _lib.check(_lib.load().comemb_build_alias(_lib.ptr(self.table), self.table_size, self.vocab_size,
                                                  _lib.ptr(self.alias), _lib.stream_ptr()))
        return self.alias

    # ---- persistence (model.py:126-140) ----------------------------------------------------------------------------------
    def save(self, path="data", file_name=None):
        if not exists(path):
            makedirs(path)
        state = {}
        for k, v in self.__dict__.items():
            state[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else v
        state.pop("_id_index", None)
        with open(path_join(path, file_name + ".bin"), "wb") as f:
            pickle.dump(state, f)

    @staticmethod
    def load_model(path="data", file_name=None, device="cuda"):
        import torch
        with open(path_join(path, file_name + ".bin"), "rb") as f:
            state = pickle.load(f)
        model = Model.__new__(Model)
        for k, v in state.items():
            if isinstance(v, np.ndarray) and k != "ground_true":
                v = torch.from_numpy(v).to(device)
            setattr(model, k, v)
        model._id_index = None
        return model
