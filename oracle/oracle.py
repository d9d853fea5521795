"""ctypes front-end of oracle/libcomemb_oracle.so and loader of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE -- may be imported only by tests/, __graft_entry__.smoke()/build() and bench.py's cpu_baseline /
`--impl reference` legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcomemb_oracle.so")

DOT_REFBLAS_QUIRK = 0  # reference as built in the authoring container (FAST_VERSION 0, OpenBLAS SkylakeX sdot)
DOT_REFBLAS = 1        # FAST_VERSION 1 flavour
DOT_WARP = 2           # summation order of the CUDA Hogwild kernels
DOT_WARP2 = 3          # ... of the size-64 specialisation (lane l owns elements 2l, 2l+1)
TOKEN_NONE = 0xFFFFFFFF

_lib = None


def build(force=False):
    src = os.path.join(HERE, "comemb_oracle.c")
    if force or not os.path.exists(LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(LIB_PATH)):
        subprocess.run(["make", "-C", HERE, "-B", "libcomemb_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        c = ctypes
        vp, i32, i64, u32, u64, f32, f64 = (c.c_void_p, c.c_int, c.c_int64, c.c_uint32, c.c_uint64, c.c_float,
                                            c.c_double)
        L.oracle_init_lut.argtypes = [vp]
        L.oracle_dot.argtypes = [vp, vp, i32, i32]
        L.oracle_dot.restype = f32
        L.oracle_train_o2.argtypes = [vp, vp, i32, vp, i64, f32, i32, i32, vp, u64, f32, u64, i32, vp]
        L.oracle_train_o2.restype = i64
        L.oracle_train_o1.argtypes = [vp, i32, u32, u32, f32, i32, vp, u64, u64, i32, vp]
        L.oracle_train_o1.restype = i64
        L.oracle_o2_walks.argtypes = [vp, vp, i32, vp, vp, i64, vp, f32, i32, i32, vp, u64, f32, i32]
        L.oracle_o2_walks.restype = i64
        L.oracle_o1_edges.argtypes = [vp, i32, vp, i64, vp, f32, i32, vp, u64, i32]
        L.oracle_o1_edges.restype = i64
        L.oracle_o3_batch.argtypes = [vp, i64, i32, vp, i64, vp, vp, vp, i32, f64, f32, i32]
        L.oracle_make_table.argtypes = [vp, i64, i64, f64, vp, i64]
        L.oracle_walk_file_seed.argtypes = [u64]
        L.oracle_walk_file_seed.restype = u64
        L.oracle_walks.argtypes = [vp, vp, i64, i32, i32, f64, u64, vp, vp]
        L.oracle_train_sg.argtypes = [vp, vp, i32, vp, i64, vp, f32, i32, i32, vp, u64, vp, vp, vp, i32, f32, f32,
                                      i32, u64, i32]
        L.oracle_train_sg.restype = i64
        L.oracle_lcg_advance.argtypes = [u64, u64]
        L.oracle_lcg_advance.restype = u64
        L.oracle_init_lut(None)
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def _chk(a, dt):
    assert isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous, (type(a), getattr(a, "dtype", None))
    return a


def init_lut():
    out = np.empty(1000, np.float32)
    lib().oracle_init_lut(_p(out))
    return out


def dot(x, y, model=DOT_REFBLAS_QUIRK):
    return float(lib().oracle_dot(_p(_chk(x, np.float32)), _p(_chk(y, np.float32)), x.size, model))


def seeds_from_numpy(rs, n):
    """The per-call seeds of pyx:427/477: (2**24)*randint(0,2**24) + randint(0,2**24), n calls in order, drawn from
    a legacy numpy RandomState (np.random module state or an explicit RandomState)."""
    r = rs.randint(0, 2 ** 24, size=(n, 2)).astype(np.uint64)
    return (r[:, 0] << np.uint64(24)) + r[:, 1]


def o2_walks(node, ctx, walks, walk_off, seeds, lr, negative, window, table, lam=1.0, dot_model=DOT_REFBLAS_QUIRK):
    """In-place: oracle of Context2Vec's single worker loop over train_o2 (pyx:454-509)."""
    _chk(node, np.float32), _chk(ctx, np.float32), _chk(walks, np.uint32), _chk(walk_off, np.int64)
    _chk(seeds, np.uint64), _chk(table, np.uint32)
    return lib().oracle_o2_walks(_p(node), _p(ctx), node.shape[1], _p(walks), _p(walk_off), len(walk_off) - 1,
                                 _p(seeds), lr, negative, window, _p(table), table.size, lam, dot_model)


def o1_edges(node, edges, seeds, lr, negative, table, dot_model=DOT_REFBLAS_QUIRK):
    """In-place: oracle of Node2Vec's single worker loop over train_o1 (pyx:407-450)."""
    _chk(node, np.float32), _chk(edges, np.uint32), _chk(seeds, np.uint64), _chk(table, np.uint32)
    return lib().oracle_o1_edges(_p(node), node.shape[1], _p(edges), edges.shape[0], _p(seeds), lr, negative,
                                 _p(table), table.size, dot_model)


def o3_batch(node, rows, mu, inv_cov, pi, beta, lr, iters=1):
    _chk(node, np.float32), _chk(rows, np.uint32), _chk(mu, np.float32), _chk(inv_cov, np.float32)
    _chk(pi, np.float32)
    lib().oracle_o3_batch(_p(node), node.shape[0], node.shape[1], _p(rows), rows.size, _p(mu), _p(inv_cov), _p(pi),
                          mu.shape[0], beta, lr, iters)


def make_table(counts, table_size, min_id=1, power=0.75):
    counts = np.ascontiguousarray(counts, np.float64)
    table = np.empty(table_size, np.uint32)
    lib().oracle_make_table(_p(counts), counts.size, min_id, power, _p(table), table_size)
    return table


def walk_file_seed(parent_seed):
    return int(lib().oracle_walk_file_seed(parent_seed))


def walks(rowptr, col, num_paths, path_length, alpha, seed):
    _chk(rowptr, np.int64), _chk(col, np.uint32)
    n = rowptr.size - 1
    out = np.empty((num_paths * n, path_length), np.uint32)
    lens = np.empty(num_paths * n, np.int32)
    lib().oracle_walks(_p(rowptr), _p(col), n, num_paths, path_length, alpha, seed, _p(out), _p(lens))
    return out, lens


def train_sg(node, negemb, path, reduced_windows, lr, negative, window, table, mu, inv_cov, pi, lambda1, lambda2,
             is_node_embedding, seed, dot_model=DOT_REFBLAS_QUIRK):
    _chk(node, np.float32), _chk(negemb, np.float32), _chk(path, np.uint32), _chk(table, np.uint32)
    if reduced_windows is not None:
        _chk(reduced_windows, np.int32)
    return lib().oracle_train_sg(_p(node), _p(negemb), node.shape[1], _p(path), path.size, _p(reduced_windows), lr,
                                 negative, window, _p(table), table.size, _p(_chk(mu, np.float32)),
                                 _p(_chk(inv_cov, np.float32)), _p(_chk(pi, np.float32)), mu.shape[0], lambda1,
                                 lambda2, is_node_embedding, seed, dot_model)


def lcg_advance(x, n):
    return int(lib().oracle_lcg_advance(x, n))


# ---- the compiled reference (oracle/_ref) ---------------------------------------------------------------------------

def ref_available(variant="tuned"):
    import sysconfig
    return os.path.exists(os.path.join(HERE, "_ref", variant, "utils",
                                       "training_sdg_inner" + sysconfig.get_config_var("EXT_SUFFIX")))


def load_ref(variant="tuned", with_python_sources=False):
    """Import the reference's compiled `utils.training_sdg_inner` from oracle/_ref/<variant>.

    with_python_sources=True additionally exposes the reference's pure-Python modules (ADSCModel.*, utils.embedding,
    utils.graph_utils ...) read in place from /root/reference -- only possible in the authoring container; used by
    tests/golden/make_golden.py, never at test/bench time on the GPU box."""
    ref_root = os.environ.get("COMEMB_REFERENCE", "/root/reference")
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(HERE, "_ref", variant, "utils")]
    if with_python_sources:
        pkg.__path__.append(os.path.join(ref_root, "utils"))
        if ref_root not in sys.path:
            sys.path.insert(0, ref_root)
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[name]
    sys.modules["utils"] = pkg
    import importlib
    return importlib.import_module("utils.training_sdg_inner")


class RefVocab(object):
    """Stand-in for utils.embedding.Vocab (embedding.py:164-175): train_o1/train_o2 only read `.index`."""
    __slots__ = ("index", "count", "sample_probability")

    def __init__(self, index, count=0):
        self.index = index
        self.count = count
        self.sample_probability = 1.0
