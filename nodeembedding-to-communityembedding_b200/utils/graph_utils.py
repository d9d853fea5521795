"""Graph ingest and random walks with the reference's names (/root/reference/utils/graph_utils.py), on the device.

The reference keeps a networkx graph and walks it in Python; here the graph is a CSR on the GPU
(`Graph.rowptr/col`, rows in node-insertion order, neighbours in adjacency order) and walks come from
csrc/walker.cu.  ORDERED mode reproduces `random.Random(seed)` walks bit-for-bit given the same row/neighbour
order as the reference's graph object; HOGWILD mode is the fast counter-based walker.
"""
import os

import numpy as np

from .. import _lib
from .._lib import MODE_HOGWILD, MODE_ORDERED, TOKEN_NONE


class Graph(object):
    """Undirected simple graph as CSR.  `ids[r]` = node id of CSR row r (first-appearance order, like
    networkx' insertion order); `col` holds CSR row numbers."""

    def __init__(self, ids, rowptr, col):
        self.ids = np.asarray(ids, np.int64)
        self.rowptr = np.ascontiguousarray(rowptr, np.int64)
        self.col = np.ascontiguousarray(col, np.uint32)
        self._dev = None

    def __len__(self):
        return self.ids.size

    def number_of_nodes(self):
        return self.ids.size

    def number_of_edges(self):
        loops = int((self.col == np.repeat(np.arange(len(self), dtype=np.uint32), np.diff(self.rowptr))).sum())
        return (self.col.size + loops) // 2

    def nodes(self):
        return self.ids.tolist()

    def degree(self):
        """{node id: degree} (what `Model(G.degree(), ...)` consumes, adsc_Karate.py:63)."""
        deg = np.diff(self.rowptr)
        return {int(i): int(d) for i, d in zip(self.ids, deg)}

    def edges(self):
        """Each undirected edge once, as id pairs, in (row, neighbour) scan order."""
        src = np.repeat(np.arange(len(self), dtype=np.int64), np.diff(self.rowptr))
        dst = self.col.astype(np.int64)
        keep = src <= dst
        # networkx reports (u, v) when v appears in u's adjacency and u was seen first
        first = np.minimum(src, dst) == src
        m = keep & first
        return np.stack([self.ids[src[m]], self.ids[dst[m]]], 1)

    def device(self):
        import torch
        if self._dev is None:
            self._dev = (torch.from_numpy(self.rowptr).cuda(), torch.from_numpy(self.col.view(np.int32)).cuda())
        return self._dev


def from_edge_array(edges, undirected=False):
    """Build the CSR from an [E,2] id array: rows in first-appearance order; adjacency in insertion order with
    duplicate edges collapsed (nx.Graph.add_edges_from semantics, graph_utils.py:61-70)."""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    flat = edges.ravel()
    uniq, first = np.unique(flat, return_index=True)
    ids = uniq[np.argsort(first, kind="stable")]
    pos = {int(v): i for i, v in enumerate(ids)}
    adj = [dict() for _ in ids]
    for u, v in edges:
        a, b = pos[int(u)], pos[int(v)]
        adj[a].setdefault(b, None)
        adj[b].setdefault(a, None)
    if undirected:
        # G.to_undirected() (graph_utils.py:106-107) copies the graph by re-inserting every adjacency entry in node
        # order (networkx >= 2), which reorders some neighbour lists; replay that pass so that walk parity holds.
        adj2 = [dict() for _ in ids]
        for a, d in enumerate(adj):
            for b in d:
                adj2[a].setdefault(b, None)
                adj2[b].setdefault(a, None)
        adj = adj2
    rowptr = np.zeros(len(ids) + 1, np.int64)
    col = []
    for i, d in enumerate(adj):
        col.extend(d.keys())
        rowptr[i + 1] = len(col)
    return Graph(ids, rowptr, np.asarray(col, np.uint32))


def from_edge_array_fast(edges, n=None):
    """Vectorised CSR build for large synthetic graphs: node ids must be 1..n, rows are in id order, neighbours sorted,
    duplicate edges and self loops dropped.  (Neighbour ORDER differs from networkx insertion order, so this builder is
    for Hogwild-mode workloads, not for ORDERED walk parity.)"""
    edges = np.asarray(edges, np.int64).reshape(-1, 2)
    if n is None:
        n = int(edges.max())
    a = np.concatenate([edges[:, 0], edges[:, 1]]) - 1
    b = np.concatenate([edges[:, 1], edges[:, 0]]) - 1
    keep = a != b
    key = np.unique(a[keep] * n + b[keep])
    src, dst = key // n, key % n
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=rowptr[1:])
    return Graph(np.arange(1, n + 1, dtype=np.int64), rowptr, dst.astype(np.uint32))


def sbm_graph(n, n_blocks, avg_degree, p_in=0.8, seed=12345):
    """Synthetic stochastic-block-model graph (BASELINE.json configs[1]): `n` nodes in `n_blocks` equal blocks, each
    node draws avg_degree/2 partners, a fraction p_in of them inside its block.  Returns (Graph, block labels)."""
    rs = np.random.RandomState(seed)
    half = max(1, avg_degree // 2)
    block = np.arange(n) % n_blocks
    src = np.repeat(np.arange(n, dtype=np.int64), half)
    intra = rs.random_sample(src.size) < p_in
    per_block = n // n_blocks
    inside = rs.randint(0, per_block, size=src.size) * n_blocks + block[src]
    anywhere = rs.randint(0, n, size=src.size)
    dst = np.where(intra, np.minimum(inside, n - 1), anywhere)
    G = from_edge_array_fast(np.stack([src + 1, dst + 1], 1), n)
    return G, block


def powerlaw_graph(n, n_edges, exponent=2.1, seed=12345, max_degree=None):
    """Chung-Lu style synthetic graph with a power-law expected-degree sequence (BASELINE.json configs[2..3] shapes:
    BlogCatalog / Youtube).  Every node gets at least one edge (a ring edge) so that no walk dead-ends."""
    rs = np.random.RandomState(seed)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** (-1.0 / (exponent - 1.0))
    if max_degree:
        w = np.minimum(w, w[0] * max_degree / (2.0 * n_edges * w[0] / w.sum()))
    p = w / w.sum()
    cdf = np.cumsum(p)
    src = np.searchsorted(cdf, rs.random_sample(n_edges)).clip(0, n - 1)
    dst = np.searchsorted(cdf, rs.random_sample(n_edges)).clip(0, n - 1)
    perm = rs.permutation(n)  # hubs scattered over the id space
    ring = np.stack([np.arange(n), (np.arange(n) + 1) % n], 1)
    e = np.concatenate([np.stack([perm[src], perm[dst]], 1), ring]) + 1
    return from_edge_array_fast(e, n)


def from_csr(ids, rowptr, col):
    return Graph(ids, rowptr, col)


def load_adjacencylist(file_, undirected=False, chunksize=10000):
    """Edge-pair text file -> Graph (graph_utils.py:72-109).  Lines starting with '#' are skipped."""
    rows = []
    with open(file_, "r") as f:
        for line in f:
            if line and line[0] != "#":
                tok = line.split()
                if tok:
                    rows.append([int(x) for x in tok[:2]])
    return from_edge_array(np.asarray(rows, np.int64), undirected=undirected)


def file_seed(rand):
    """The per-file generator seed of graph_utils.py:150: rand.randint(0, 2**31)."""
    return rand.randint(0, 2 ** 31)


def build_deepwalk_corpus(G, num_paths, path_length, alpha=0, rand=None, seed=None, mode=MODE_ORDERED,
                          return_device=False, first_walk=0, n_out=None):
    """Walks as CSR-row tokens: uint32 [num_paths*len(G), path_length] padded with TOKEN_NONE, and their lengths
    (graph_utils.py:172-197).  ORDERED needs an int `seed` = the seed of the reference's random.Random."""
    import torch
    if seed is None:
        if rand is None:
            raise ValueError("pass seed= (int) or rand= (random.Random) to derive one")
        seed = file_seed(rand)
    n = len(G)
    total = num_paths * n
    if n_out is None:
        n_out = total - first_walk
    rowptr, col = G.device()
    walks = torch.empty((n_out, path_length), dtype=torch.int32, device="cuda")
    lens = torch.empty(n_out, dtype=torch.int32, device="cuda")
    st = _lib.load().comemb_walks_csr(_lib.ptr(rowptr), _lib.ptr(col), n, int(num_paths), int(path_length),
                                      float(alpha), int(seed), int(mode), int(first_walk), int(n_out),
                                      _lib.ptr(walks), _lib.ptr(lens), _lib.stream_ptr())
    _lib.check(st)
    if return_device:
        return walks, lens
    return walks.cpu().numpy().view(np.uint32), lens.cpu().numpy()


def build_deepwalk_corpus_iter(G, num_paths, path_length, alpha=0, rand=None, seed=None, mode=MODE_ORDERED):
    """Generator of walks as lists of node ids (graph_utils.py:191-197)."""
    walks, lens = build_deepwalk_corpus(G, num_paths, path_length, alpha, rand, seed, mode)
    for w, l in zip(walks, lens):
        yield G.ids[w[:l].astype(np.int64)].tolist()


def write_walks_to_disk(G, filebase, num_paths, path_length, alpha=0, rand=None, num_workers=1, mode=MODE_ORDERED):
    """One text file per worker, one walk per line, space-separated node ids (graph_utils.py:122-156).  Files get
    seeds rand.randint(0, 2**31) in order, passes are split over files like the reference does."""
    import random as _random
    rand = rand or _random.Random(0)
    if num_paths <= num_workers:
        per_file = [1] * num_paths
    else:
        step = int(num_paths / num_workers) + 1
        per_file = [len(range(s, min(num_paths, s + step))) for s in range(0, num_paths, step)]
    files = []
    os.makedirs(os.path.dirname(filebase) or ".", exist_ok=True)
    for x, ppw in enumerate(per_file):
        fname = "{}.{}".format(filebase, x)
        seed = file_seed(rand)
        with open(fname, "w") as fout:
            for walk in build_deepwalk_corpus_iter(G, ppw, path_length, alpha=alpha, seed=seed, mode=mode):
                fout.write(u"{}\n".format(u" ".join(str(v) for v in walk)))
        files.append(fname)
    return files


def combine_files_iter(file_list):
    for file in file_list:
        if os.path.isfile(file):
            with open(file, "r") as f:
                for line in f:
                    yield np.array([int(node) for node in line.split()])


def count_lines(f):
    return sum(1 for _ in open(f)) if os.path.isfile(f) else 0
