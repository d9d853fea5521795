"""Drop-in for the reference's `utils.training_sdg_inner` (Cython, /root/reference/utils/training_sdg_inner.pyx).

Same names, argument meaning and return values:
    train_o1 (pyx:407)   train_o2 (pyx:454)   train_sg (stale .c:2736 / utils/embedding.py:15)   init (pyx:512)
    FAST_VERSION (pyx:549)   REAL (pyx:15)
plus the batch entry points the learners use (one launch per edge list / walk corpus instead of one call per item):
    o1_batch, o2_batch, sg_batch, o3_batch.

Tables may be numpy arrays (the reference's convention: borrowed, updated IN PLACE -- here that means host->device,
kernel, device->host inside the call) or torch CUDA tensors (device resident, no copies).  All arithmetic runs in
csrc/*.cu through the C ABI; there is no CPU fallback.
"""
import numpy as np

from .. import _lib
from .._lib import (F_ALIAS, F_ATOMIC, F_DOT_FLOAT, F_SEED_HASH, MODE_HOGWILD, MODE_ORDERED, TOKEN_NONE,  # noqa: F401
                    ComembError)

REAL = np.float32  # pyx:15
MAX_SENTENCE_LEN = 10000  # pyx:18


def init():
    """Build the sigmoid table and upload it (pyx:512-549).  Returns 0: ORDERED mode models the reference's
    FAST_VERSION-0 arithmetic by default (pass flags=F_DOT_FLOAT for the FAST_VERSION-1 flavour)."""
    return _lib.ensure_init()


def __getattr__(name):  # FAST_VERSION = init() at first use (the reference runs init() at import, pyx:549)
    if name == "FAST_VERSION":
        return init()
    raise AttributeError(name)


# ---- host <-> device plumbing ----------------------------------------------------------------------------------------
_const_cache = {}  # (host address, nbytes, dtype, crc32 of the contents) -> device copy of a negative table


def clear_cache():
    _const_cache.clear()


def _torch():
    import torch
    return torch


class _Borrowed(object):
    """A table argument: torch CUDA tensor used as is, or a numpy array mirrored on the device and written back."""

    def __init__(self, arr, dtype=np.float32, writeback=True):
        torch = _torch()
        self.host = None
        if isinstance(arr, torch.Tensor):
            if not arr.is_cuda:
                raise ComembError("torch tensors passed to the ComEmb kernels must live on a CUDA device")
            if arr.dtype != getattr(torch, np.dtype(dtype).name) or not arr.is_contiguous():
                raise ComembError("expected a contiguous %s tensor, got %s" % (np.dtype(dtype).name, arr.dtype))
            self.dev = arr
        else:
            a = np.asarray(arr)
            if a.dtype != dtype or not a.flags.c_contiguous:
                # the reference reinterprets the raw buffer (pyx:410-411) -- garbage in, garbage out; we refuse instead
                raise ComembError("expected a C-contiguous %s array, got %s" % (np.dtype(dtype).name, a.dtype))
            self.dev = torch.from_numpy(a).cuda()
            if writeback:
                self.host = a

    def writeback(self):
        if self.host is not None:
            self.host[...] = self.dev.cpu().numpy()


def _const_dev(arr, dtype, cache=False):
    """Read-only input as a device tensor.  numpy inputs are uploaded on every call; only the negative table
    (`cache=True`: tens to hundreds of MB, rebuilt rarely) is kept on the device between calls, keyed by its address,
    size AND a checksum of its full contents, so an in-place rebuild can never be served from a stale copy.  The small
    GMM inputs (centroid / inv_cov / pi, refit in place between calls) are never cached."""
    torch = _torch()
    if arr is None:
        return None
    if isinstance(arr, torch.Tensor):
        want = {np.dtype(np.uint32): torch.uint32}.get(np.dtype(dtype), getattr(torch, np.dtype(dtype).name, None))
        if arr.dtype == torch.int32 and np.dtype(dtype) == np.uint32:
            return arr.contiguous()
        if arr.dtype != want:
            raise ComembError("expected %s, got %s" % (want, arr.dtype))
        return arr.contiguous()
    a = np.ascontiguousarray(arr, dtype=dtype)
    view = a.view(np.int32) if a.dtype == np.uint32 else a
    if not cache:
        return torch.from_numpy(view).cuda()
    import zlib
    key = (a.ctypes.data, a.nbytes, a.dtype.str, zlib.crc32(memoryview(a).cast("B")))
    hit = _const_cache.get(key)
    if hit is None:
        if len(_const_cache) > 4:
            _const_cache.clear()
        hit = torch.from_numpy(view).cuda()
        _const_cache[key] = hit
    return hit


def _dev(a, np_dtype):
    torch = _torch()
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.contiguous() if a.is_cuda else a.contiguous().cuda()
    a = np.ascontiguousarray(a, dtype=np_dtype)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).cuda()


def draw_seeds(n, rs=None):
    """The per-call LCG seeds of pyx:427 / pyx:477: `(2**24)*randint(0,2**24) + randint(0,2**24)` for n calls, in
    call order, from the legacy global `np.random` state (or an explicit RandomState)."""
    rs = np.random if rs is None else rs
    r = rs.randint(0, 2 ** 24, size=(n, 2)).astype(np.uint64)
    return (r[:, 0] << np.uint64(24)) + r[:, 1]


def _indices(py_items):
    """list of Vocab/None -> uint32 row indices with TOKEN_NONE for None (pyx:483-490)."""
    return np.fromiter((TOKEN_NONE if w is None else w.index for w in py_items), dtype=np.uint32,
                       count=len(py_items))


# ---- batch entry points -------------------------------------------------------------------------------------------------
def check_row_tokens(tokens, n_rows, what="walk tokens"):
    """Raise if a uint32 row token (other than TOKEN_NONE) addresses a row >= n_rows.  The kernels, like the
    reference (pyx:410-411), do not bounds-check; the learners call this once per corpus so that a bad id becomes a
    Python error instead of an illegal device access."""
    torch = _torch()
    if tokens is None or tokens.numel() == 0:
        return
    t = tokens.reshape(-1)
    t64 = t.to(torch.int64) & 0xFFFFFFFF  # stored as int32 bit patterns
    bad = (t64 >= int(n_rows)) & (t64 != TOKEN_NONE)
    if bool(bad.any()):
        raise ComembError("%s: row index %d out of range for a table of %d rows" % (what, int(t64[bad][0]), n_rows))


def hogwild_concurrency(n_rows, workers=1):
    """Cap on concurrently processed walks/edges in HOGWILD mode: the reference's `workers`, or as many as keep the
    expected number of in-flight updates per table row below ~1/4 (7 rows per pair -> n_rows/28), whichever is larger.
    100K rows -> 3571 (a full B200 holds 3552 warps of the o2 kernel); karate -> `workers`."""
    return int(max(int(workers), int(n_rows) // 28, 1))


# The ORDERED replay is bound by the dependency chain of the update stream, not by the GPU: on the walk corpora of
# BASELINE.json consecutive walks share ~15 rows, the longest dependency path covers 68-91 % of all pairs (DESIGN.md
# section 6, scripts/ordered_critical_path.py), and the replay runs at ~1.4e6 pair updates/s.  A learner built with
# workers=1 and no explicit mode therefore replays the reference's order only up to this many updates (about a second
# and a half) and trains larger corpora lock-free like the reference's workers>1; mode="ordered" always replays.
ORDERED_AUTO_MAX_UPDATES = 2000000


def select_mode(mode, workers, n_updates):
    """The learners' mode rule.  mode: None | "ordered" | "hogwild" | MODE_*; n_updates: upper bound of the pair
    updates of the call (o2: tokens * 2 * window; o1: 2 * edges)."""
    if mode is not None:
        return {"ordered": MODE_ORDERED, "hogwild": MODE_HOGWILD}.get(mode, mode)
    if workers == 1 and n_updates <= ORDERED_AUTO_MAX_UPDATES:
        return MODE_ORDERED
    if workers == 1:
        import logging
        logging.getLogger(__name__).info(
            "workers=1 with %d pair updates (> %d): training lock-free; pass mode=\"ordered\" for the reference's "
            "sequential order bit for bit (~1.4e6 updates/s)", n_updates, ORDERED_AUTO_MAX_UPDATES)
    return MODE_HOGWILD


class _max_warps(object):
    """Concurrency cap for the launches inside the block; the caller's previous cap is restored afterwards."""

    def __init__(self, n):
        self.n = n
        self.ctx = None

    def __enter__(self):
        if self.n:
            self.ctx = _lib.opts(max_warps=int(self.n))
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def o2_batch(node, ctx, walks, walk_off, seeds, lr, negative, window, table, alpha=1.0, mode=MODE_ORDERED, flags=0,
             alias=None, base_seed=0, count_tokens=False, max_warps=0):
    """train_o2 over a corpus.  node/ctx: float32 CUDA tensors [N, d] (in place).  walks: uint32 row tokens (flat),
    walk_off: int64 [n_walks+1], seeds: uint64 [n_walks] or None (-> F_SEED_HASH from base_seed)."""
    torch = _torch()
    _lib.ensure_init()
    n_walks = int(walk_off.numel()) - 1
    if n_walks <= 0 or walks.numel() == 0:  # an empty corpus (or only empty walks): nothing to update, 0 tokens
        return 0 if count_tokens else None
    if seeds is None:
        flags |= F_SEED_HASH
    if alias is not None:
        flags |= F_ALIAS
    tok = torch.zeros(1, dtype=torch.int64, device=node.device) if count_tokens else None
    with _max_warps(max_warps):
        st = _lib.load().comemb_o2_walks(
            _lib.ptr(node), _lib.ptr(ctx), node.shape[0], node.shape[1], _lib.ptr(walks), _lib.ptr(walk_off), n_walks,
            _lib.ptr(seeds), int(base_seed), _lib.ptr(table), 0 if table is None else table.numel(), _lib.ptr(alias),
            0 if alias is None else alias.numel() // 2, int(window), int(negative), float(lr), float(alpha),
            int(mode), int(flags), _lib.ptr(tok), _lib.stream_ptr())
    _lib.check(st)
    return int(tok.item()) if count_tokens else None


class HostO2Runner(object):
    """train_o2 over a corpus whose tables and walks live in HOST memory (the reference's calling convention at batch
    granularity): every call copies node/ctx/walks/seeds host->device from pinned staging buffers, runs the kernel
    and copies both tables back, all on the current stream.  Used for end-to-end measurements."""

    def __init__(self, n_rows, size, max_tokens, max_walks, table):
        torch = _torch()
        self.table = _const_dev(table, np.uint32, cache=True)
        pin = dict(pin_memory=True)
        self.h_node = torch.empty((n_rows, size), dtype=torch.float32, **pin)
        self.h_ctx = torch.empty((n_rows, size), dtype=torch.float32, **pin)
        self.h_walks = torch.empty(max_tokens, dtype=torch.int32, **pin)
        self.h_off = torch.empty(max_walks + 1, dtype=torch.int64, **pin)
        self.h_seeds = torch.empty(max_walks, dtype=torch.int64, **pin)
        self.d_node = torch.empty((n_rows, size), dtype=torch.float32, device="cuda")
        self.d_ctx = torch.empty_like(self.d_node)
        self.d_walks = torch.empty(max_tokens, dtype=torch.int32, device="cuda")
        self.d_off = torch.empty(max_walks + 1, dtype=torch.int64, device="cuda")
        self.d_seeds = torch.empty(max_walks, dtype=torch.int64, device="cuda")

    def host_tables(self):
        """numpy views of the runner's page-locked table buffers: a caller that keeps its tables here avoids the
        pageable->pinned staging copy (pass these same arrays to run())."""
        return self.h_node.numpy(), self.h_ctx.numpy()

    def host_walk_buffers(self, n_tokens, n_walks):
        """numpy views (walks uint32 [n_tokens], walk_off int64 [n_walks+1], seeds uint64 [n_walks]) of the runner's
        page-locked input buffers: a producer that writes its walks here and passes these views to run() avoids the
        pageable->pinned staging copy, like host_tables() for the tables."""
        return (self.h_walks.numpy()[:n_tokens].view(np.uint32), self.h_off.numpy()[:n_walks + 1],
                self.h_seeds.numpy()[:n_walks].view(np.uint64))

    def run(self, node, ctx, walks, walk_off, seeds, lr, negative, window, alpha=1.0, mode=MODE_HOGWILD, flags=0):
        """node/ctx: numpy float32 [N,d] updated in place; walks uint32 flat; walk_off int64; seeds uint64.
        Returns (h2d_bytes, d2h_bytes)."""
        torch = _torch()
        nt, nw = walks.size, walk_off.size - 1
        own = node is self.h_node.numpy() or (node.ctypes.data == self.h_node.data_ptr() and
                                              ctx.ctypes.data == self.h_ctx.data_ptr())
        if not own:
            self.h_node.numpy()[...] = node
            self.h_ctx.numpy()[...] = ctx
        if walks.ctypes.data != self.h_walks.data_ptr():
            self.h_walks.numpy()[:nt] = walks.view(np.int32)
        if walk_off.ctypes.data != self.h_off.data_ptr():
            self.h_off.numpy()[:nw + 1] = walk_off
        if seeds.ctypes.data != self.h_seeds.data_ptr():
            self.h_seeds.numpy()[:nw] = seeds.view(np.int64)
        self.d_node.copy_(self.h_node, non_blocking=True)
        self.d_ctx.copy_(self.h_ctx, non_blocking=True)
        self.d_walks[:nt].copy_(self.h_walks[:nt], non_blocking=True)
        self.d_off[:nw + 1].copy_(self.h_off[:nw + 1], non_blocking=True)
        self.d_seeds[:nw].copy_(self.h_seeds[:nw], non_blocking=True)
        o2_batch(self.d_node, self.d_ctx, self.d_walks[:nt], self.d_off[:nw + 1], self.d_seeds[:nw], lr, negative,
                 window, self.table, alpha=alpha, mode=mode, flags=flags)
        self.h_node.copy_(self.d_node, non_blocking=True)
        self.h_ctx.copy_(self.d_ctx, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if not own:
            node[...] = self.h_node.numpy()
            ctx[...] = self.h_ctx.numpy()
        h2d = 2 * node.nbytes + nt * 4 + (nw + 1) * 8 + nw * 8
        return h2d, 2 * node.nbytes


def o1_batch(node, edges, seeds, lr, negative, table, mode=MODE_ORDERED, flags=0, alias=None, base_seed=0,
             edge_stride=0, max_warps=0):
    """train_o1 over an edge list.  edges: uint32 CUDA tensor [E, 2] of row indices."""
    _lib.ensure_init()
    if seeds is None:
        flags |= F_SEED_HASH
    if alias is not None:
        flags |= F_ALIAS
    n_edges = int(edges.numel() // 2)
    with _max_warps(max_warps):
        st = _lib.load().comemb_o1_edges(
            _lib.ptr(node), node.shape[0], node.shape[1], _lib.ptr(edges), n_edges, _lib.ptr(seeds), int(base_seed),
            _lib.ptr(table), 0 if table is None else table.numel(), _lib.ptr(alias),
            0 if alias is None else alias.numel() // 2, int(negative), float(lr), int(mode), int(flags),
            int(edge_stride), _lib.stream_ptr())
        _lib.check(st)


def transpose_blocks(inv_cov):
    torch = _torch()
    out = torch.empty_like(inv_cov)
    _lib.check(_lib.load().comemb_transpose_blocks(_lib.ptr(inv_cov), _lib.ptr(out), inv_cov.shape[0],
                                                   inv_cov.shape[1], _lib.stream_ptr()))
    return out


def o3_batch(node, rows, centroid, inv_cov_t, pi, beta, lr, iters=1):
    """Community2Vec.train's update (community_embeddings.py:61-77) on CUDA tensors; inv_cov_t from
    transpose_blocks().  rows: uint32 CUDA tensor of selected rows or None for all."""
    st = _lib.load().comemb_o3_batch(
        _lib.ptr(node), node.shape[0], node.shape[1], _lib.ptr(rows), 0 if rows is None else rows.numel(),
        _lib.ptr(centroid), _lib.ptr(inv_cov_t), _lib.ptr(pi), centroid.shape[0], float(beta), float(lr), int(iters),
        _lib.stream_ptr())
    _lib.check(st)


def pi_top1(pi):
    """Dense responsibilities [N,K] -> (comm int32 [N], weight float32 [N]): the single non-zero community of each row
    (-1 / 0.0 for all-zero rows).  Raises if a row has more than one non-zero entry."""
    torch = _torch()
    nz = pi != 0
    cnt = nz.sum(1)
    if bool((cnt > 1).any()):
        raise ComembError("pi is not one-hot: %d rows have several non-zero responsibilities" % int((cnt > 1).sum()))
    comm = torch.where(cnt == 1, nz.float().argmax(1), torch.full_like(cnt, -1)).to(torch.int32)
    weight = pi.sum(1).to(torch.float32)
    return comm.contiguous(), weight.contiguous()


def o3_batch_top1(node, rows, centroid, inv_cov_t, comm, weight, beta, lr, iters=1):
    """o3_batch with pi in top-1 form (see pi_top1)."""
    st = _lib.load().comemb_o3_batch_top1(
        _lib.ptr(node), node.shape[0], node.shape[1], _lib.ptr(rows), 0 if rows is None else rows.numel(),
        _lib.ptr(centroid), _lib.ptr(inv_cov_t), _lib.ptr(comm), _lib.ptr(weight), centroid.shape[0], float(beta),
        float(lr), int(iters), _lib.stream_ptr())
    _lib.check(st)


def sg_batch_top1(node, negemb, walks, walk_off, reduced_windows, seeds, lr, negative, window, table, centroid, inv_cov,
                  comm, weight, lambda1, lambda2, flags=0, base_seed=0):
    """The fused pass (Hogwild, tensor-core kernel) with pi in top-1 form: no dense [N,K] matrix is needed."""
    _lib.ensure_init()
    if seeds is None:
        flags |= F_SEED_HASH
    st = _lib.load().comemb_sg_fused_top1(
        _lib.ptr(node), _lib.ptr(negemb), node.shape[0], node.shape[1], _lib.ptr(walks), _lib.ptr(walk_off),
        int(walk_off.numel()) - 1, _lib.ptr(reduced_windows), _lib.ptr(seeds), int(base_seed), _lib.ptr(table),
        table.numel(), _lib.ptr(centroid), _lib.ptr(inv_cov), _lib.ptr(comm), _lib.ptr(weight), centroid.shape[0],
        int(window), int(negative), float(lr), float(lambda1), float(lambda2), int(flags), _lib.stream_ptr())
    _lib.check(st)


def sg_batch(node, negemb, walks, walk_off, reduced_windows, seeds, lr, negative, window, table, centroid, inv_cov, pi,
             lambda1, lambda2, is_node_embedding, mode=MODE_ORDERED, flags=0, base_seed=0):
    """The legacy fused pass (stale train_sg) over a corpus."""
    _lib.ensure_init()
    if seeds is None:
        flags |= F_SEED_HASH
    n_walks = int(walk_off.numel()) - 1
    K = 0 if centroid is None else centroid.shape[0]
    st = _lib.load().comemb_sg_fused(
        _lib.ptr(node), _lib.ptr(negemb), node.shape[0], node.shape[1], _lib.ptr(walks), _lib.ptr(walk_off), n_walks,
        _lib.ptr(reduced_windows), _lib.ptr(seeds), int(base_seed), _lib.ptr(table),
        0 if table is None else table.numel(), _lib.ptr(centroid), _lib.ptr(inv_cov), _lib.ptr(pi), K, int(window),
        int(negative), float(lr), float(lambda1), float(lambda2), int(is_node_embedding), int(mode), int(flags),
        _lib.stream_ptr())
    _lib.check(st)


def o2_pos_loss(node, ctx, walks, walk_off, window):
    """(sum over window pairs of -log sigmoid(x_j . c_i), number of pairs) -- evaluation helper."""
    torch = _torch()
    out = torch.zeros(2, dtype=torch.float64, device=node.device)
    _lib.check(_lib.load().comemb_o2_pos_loss(_lib.ptr(node), _lib.ptr(ctx), node.shape[1], _lib.ptr(walks),
                                              _lib.ptr(walk_off), int(walk_off.numel()) - 1, int(window),
                                              _lib.ptr(out), _lib.stream_ptr()))
    s, n = out.tolist()
    return s, int(n)


# ---- the reference's per-call entry points ----------------------------------------------------------------------------
def train_o1(py_node_embedding, py_edge, py_lr, py_negative, py_table, py_size=None, py_work=None, flags=0):
    """pyx:407-450.  One edge = two directed first-order updates sharing one LCG stream.  Returns the number of
    non-None entries of `py_edge` (pyx:439).  `py_work` is accepted for signature compatibility (scratch lives on
    the device); `py_size` must equal the row length as in the reference."""
    torch = _torch()
    seed = draw_seeds(1)  # pyx:427 -- drawn before anything else, like the reference
    node = _Borrowed(py_node_embedding)
    _check_size(node.dev, py_size)
    idx = _indices(py_edge[:MAX_SENTENCE_LEN])
    result = int((idx != TOKEN_NONE).sum())
    if len(idx) < 2 or (idx[:2] == TOKEN_NONE).any():
        # the reference reads uninitialised indexes[] here (pyx:444); we refuse
        raise ComembError("train_o1 needs an edge of two in-vocabulary nodes")
    edges = torch.from_numpy(idx[:2].view(np.int32).copy()).cuda()
    o1_batch(node.dev, edges, _dev(seed, np.uint64), py_lr, py_negative, _const_dev(py_table, np.uint32, cache=True),
             mode=MODE_ORDERED, flags=flags)
    node.writeback()
    return result


def train_o2(py_node_embedding, py_context_embedding, py_path, py_lr, py_negative, py_window, py_table,
             py_alpha=1.0, py_size=None, py_work=None, flags=0):
    """pyx:454-509.  One walk: every (centre i, neighbour j) pair inside the window is one SGNS update of
    node[path[j]] against context[path[i]] and `py_negative` table draws.  Returns the number of non-None tokens."""
    torch = _torch()
    seed = draw_seeds(1)  # pyx:477
    node = _Borrowed(py_node_embedding)
    ctx = _Borrowed(py_context_embedding)
    _check_size(node.dev, py_size)
    idx = _indices(py_path[:MAX_SENTENCE_LEN])  # pyx:480
    result = int((idx != TOKEN_NONE).sum())
    if len(idx):
        walks = torch.from_numpy(idx.view(np.int32)).cuda()
        off = torch.tensor([0, len(idx)], dtype=torch.int64, device="cuda")
        o2_batch(node.dev, ctx.dev, walks, off, _dev(seed, np.uint64), py_lr, py_negative, py_window,
                 _const_dev(py_table, np.uint32, cache=True), alpha=py_alpha, mode=MODE_ORDERED, flags=flags)
    node.writeback()
    ctx.writeback()
    return result


def train_sg(py_node_embedding, py_negative_embedding, py_path, py_alpha, py_negative, py_window, py_table,
             py_centroid, py_inv_covariance_mat, py_pi, py_k, py_covariance_mat, py_lambda1=1.0, py_lambda2=0.0,
             py_size=None, py_work=None, py_work_o3=None, py_work1_o3=None, py_work2_o3=None,
             py_is_node_embedding=1, flags=0):
    """The legacy fused pass (stale utils/training_sdg_inner.c:2736; signature of utils/embedding.py:15-18):
    per pair, the o3 gradient of x_j, the SGNS update and the combined write.  np.random draw order as the compiled
    original: the LCG seed (2 draws) first, then one randint(window) per token when window > 1 (c:3194, c:3364)."""
    torch = _torch()
    seed = draw_seeds(1)
    node = _Borrowed(py_node_embedding)
    same = py_negative_embedding is py_node_embedding
    neg = node if same else _Borrowed(py_negative_embedding)
    _check_size(node.dev, py_size)
    idx = _indices(py_path[:MAX_SENTENCE_LEN])
    result = int((idx != TOKEN_NONE).sum())
    rw = None
    if py_window > 1:
        rw = np.array([np.random.randint(py_window) if t != TOKEN_NONE else 0 for t in idx], dtype=np.int32)
    if len(idx):
        walks = torch.from_numpy(idx.view(np.int32)).cuda()
        off = torch.tensor([0, len(idx)], dtype=torch.int64, device="cuda")
        sg_batch(node.dev, neg.dev, walks, off, _dev(rw, np.int32), _dev(seed, np.uint64), py_alpha, py_negative,
                 py_window, _const_dev(py_table, np.uint32, cache=True), _const_dev(py_centroid, np.float32),
                 _const_dev(py_inv_covariance_mat, np.float32), _const_dev(py_pi, np.float32), py_lambda1, py_lambda2,
                 py_is_node_embedding, mode=MODE_ORDERED, flags=flags)
    node.writeback()
    if not same:
        neg.writeback()
    return result


def train_sg_twin(py_node_embedding, py_context_embedding, py_path, py_alpha, py_negative, py_window, py_table,
                  py_centroid, py_inv_covariance_mat, py_pi, py_k, py_covariance_mat, py_lambda1=1.0, py_lambda2=0.0,
                  py_size=None, py_work=None, py_work_o3=None, py_work1_o3=None, py_work2_o3=None,
                  py_is_node_embedding=1):
    """The reference's pure-Python fallback `train_sg` (utils/embedding.py:15-72) -- the semantics it falls back to
    whenever the compiled fused kernel is absent, i.e. always at HEAD: no window shrinking, negatives redrawn with
    `np.random.randint(len(table))` until they differ from both nodes of the pair (same draw order, so a seeded run
    picks the same targets), exact sigmoid, vectorised target update.  The host walks the path and draws the
    targets; the arithmetic runs in csrc (comemb_sg_twin)."""
    torch = _torch()
    node = _Borrowed(py_node_embedding)
    same = py_context_embedding is py_node_embedding
    ctx = node if same else _Borrowed(py_context_embedding)
    table = np.asarray(py_table.cpu().numpy() if isinstance(py_table, torch.Tensor) else py_table)
    rows, targets = [], []
    n_table = table.shape[0]
    for pos, nd in enumerate(py_path):  # embedding.py:29-52
        if nd is None:
            continue
        start = max(0, pos - py_window)
        for pos2, nd2 in enumerate(py_path[start: pos + py_window + 1], start):
            if nd2 and not (pos2 == pos):
                idx = [nd.index]
                while len(idx) < py_negative + 1:
                    w = int(table[np.random.randint(n_table)])
                    if w != nd.index and w != nd2.index:
                        idx.append(w)
                rows.append(nd2.index)
                targets.append(idx)
    if rows:
        # keep every device temporary referenced until the kernel has run
        d_rows = _dev(np.asarray(rows, np.uint32), np.uint32)
        d_tgt = _dev(np.asarray(targets, np.uint32).reshape(-1), np.uint32)
        d_mu, d_inv = _const_dev(py_centroid, np.float32), _const_dev(py_inv_covariance_mat, np.float32)
        d_pi = _const_dev(py_pi, np.float32)
        st = _lib.load().comemb_sg_twin(
            _lib.ptr(node.dev), _lib.ptr(ctx.dev), node.dev.shape[0], node.dev.shape[1], _lib.ptr(d_rows),
            _lib.ptr(d_tgt), len(rows), int(py_negative), float(py_alpha), float(py_lambda1), float(py_lambda2),
            _lib.ptr(d_mu), _lib.ptr(d_inv), _lib.ptr(d_pi), int(py_k), int(py_is_node_embedding), _lib.stream_ptr())
        _lib.check(st)
        torch.cuda.current_stream().synchronize()
        del d_rows, d_tgt
    node.writeback()
    if not same:
        ctx.writeback()
    return len([w for w in py_path if w is not None])


def _check_size(dev, py_size):
    if dev.dim() != 2:
        raise ComembError("embedding tables must be 2-D [n_rows, size]")
    if py_size is not None and int(py_size) != dev.shape[1]:
        raise ComembError("py_size=%s does not match the row length %d" % (py_size, dev.shape[1]))
