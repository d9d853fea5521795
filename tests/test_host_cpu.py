"""CPU-only tests (no CUDA call): the C-ABI library loads and exports every declared symbol, the host-side mirror of
the reference interface behaves like the reference's Python, and the product never touches oracle/."""
import os
import random
import re

import numpy as np
import pytest

import cases
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nodeembedding-to-communityembedding_b200")


def test_capi_exports_every_declared_symbol():
    from comemb_b200 import _lib
    header = open(os.path.join(ROOT, "include", "comemb_b200.h")).read()
    declared = set(re.findall(r"\b(comemb_[a-z0-9_]+)\s*\(", header))
    assert {"comemb_init", "comemb_o2_walks", "comemb_o1_edges", "comemb_o3_batch", "comemb_sg_fused",
            "comemb_walks_csr", "comemb_make_table"} <= declared
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.comemb_abi_version() == 2
    assert b"invalid argument" in lib.comemb_error_string(-1)


def test_launch_options_are_per_thread_and_restored():
    """comemb_opts_t lives in thread-local storage (no CUDA call involved): a worker thread's settings never leak into
    another thread, `_lib.opts(...)` restores what was there before, and invalid values are refused."""
    import ctypes
    import threading
    from comemb_b200 import _lib
    lib = _lib.load()
    assert lib.comemb_set_opts(None) == 0
    with _lib.opts(max_warps=7):
        assert _lib.get_opts().max_warps == 7
        seen = {}

        def other():
            seen["before"] = _lib.get_opts().max_warps
            with _lib.opts(max_warps=3, variant=_lib.VARIANT_GENERIC):
                seen["inside"] = (_lib.get_opts().max_warps, _lib.get_opts().variant)
        t = threading.Thread(target=other)
        t.start()
        t.join()
        assert seen == {"before": 0, "inside": (3, 9)}
        with _lib.opts(variant=_lib.VARIANT_ROUND1):
            o = _lib.get_opts()
            assert (o.max_warps, o.variant) == (7, 6)
        assert _lib.get_opts().variant == 0 and _lib.get_opts().max_warps == 7
    assert _lib.get_opts().max_warps == 0
    assert lib.comemb_set_tuning(4, 80, 903) == 0
    o = _lib.get_opts()
    assert (o.centres_per_unit, o.max_walk_len, o.blocks_per_sm, o.variant) == (4, 80, 3, 9)
    bad = _lib.ComembOpts(0, 0, 0, 0, -1)
    assert lib.comemb_set_opts(ctypes.byref(bad)) == -1
    assert lib.comemb_set_opts(None) == 0 and _lib.get_opts().variant == 0


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libcomemb_oracle" not in src, f


def test_missing_library_fails_loudly(monkeypatch):
    from comemb_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcomemb_b200.so")
    with pytest.raises(_lib.ComembError):
        _lib.load()


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    import comemb_b200.utils.training_sdg_inner as K
    node = np.zeros((4, 8), np.float32)
    with pytest.raises(Exception):
        K.train_o2(node, node.copy(), [O.RefVocab(0), O.RefVocab(1)], 0.1, 1, 1, np.ones(4, np.uint32), py_size=8)


def test_draw_seeds_equals_the_reference_expression():
    import comemb_b200.utils.training_sdg_inner as K
    np.random.seed(5)
    want = [(2 ** 24) * np.random.randint(0, 2 ** 24) + np.random.randint(0, 2 ** 24) for _ in range(50)]  # pyx:477
    np.random.seed(5)
    got = K.draw_seeds(50)
    assert got.dtype == np.uint64 and got.tolist() == want


class _M(object):
    pass


def _model(ids, probs=None):
    from comemb_b200.utils.embedding import Vocab
    m = _M()
    m.vocab = {}
    for i, node in enumerate(sorted(ids)):
        m.vocab[node] = Vocab(index=i, count=1, sample_probability=1.0 if probs is None else probs[i])
    from comemb_b200.ADSCModel.model import Model
    m._id_index = None
    m.id_index = lambda: Model.id_index(m)
    return m


def test_paths_to_rows_matches_prepare_sentences():
    from comemb_b200.utils.embedding import paths_to_rows, prepare_sentences, chunkize_serial, RepeatCorpusNTimes
    m = _model([1, 2, 3, 5, 8, 13])
    paths = [np.array([1, 2, 99, 5]), [13, 13, 4], [], np.array([8])]
    flat, off = paths_to_rows(m, paths)
    ref = [[v.index for v in p] for p in prepare_sentences(m, paths)]
    assert [flat[off[i]:off[i + 1]].tolist() for i in range(len(paths))] == ref == [[0, 1, 3], [5, 5], [], [4]]
    # down-sampling consumes np.random.random_sample() in the same order
    probs = [1.0, 0.5, 1.0, 0.2, 1.0, 0.7]
    m = _model([1, 2, 3, 5, 8, 13], probs)
    big = [np.random.RandomState(1).choice([1, 2, 3, 5, 8, 13, 77], size=200)]
    np.random.seed(3)
    ref = [[v.index for v in p] for p in prepare_sentences(m, big)]
    np.random.seed(3)
    flat, off = paths_to_rows(m, big)
    assert flat.tolist() == ref[0]
    # rectangular corpora take the vectorised path: same tokens, same np.random consumption
    rect = np.random.RandomState(2).choice([1, 2, 3, 5, 8, 13, 77], size=(50, 6))
    np.random.seed(4)
    ref = [[v.index for v in p] for p in prepare_sentences(m, rect)]
    np.random.seed(4)
    flat, off = paths_to_rows(m, rect)
    assert [flat[off[i]:off[i + 1]].tolist() for i in range(50)] == ref
    np.random.seed(4)
    flat2, off2 = paths_to_rows(m, RepeatCorpusNTimes(rect, 1))
    assert np.array_equal(flat, flat2) and np.array_equal(off, off2)
    assert list(chunkize_serial(range(7), 3)) == [[0, 1, 2], [3, 4, 5], [6]]
    assert list(RepeatCorpusNTimes([1, 2], 2)) == [1, 2, 1, 2]


def test_graph_from_karate_file_reproduces_the_reference_graph(golden):
    """Row order, neighbour order and edge order of the networkx graph the reference builds (golden CSR extracted
    from its own graph object, make_golden.py) -- needed for bit-exact ORDERED walks from a file."""
    import comemb_b200.utils.graph_utils as gu
    G = gu.load_adjacencylist(os.path.join(ROOT, "tests", "golden", "karate.adjlist"), True)
    g = golden["walks"]
    assert np.array_equal(G.ids, g["karate/ids"])
    assert np.array_equal(G.rowptr, g["karate/rowptr"])
    assert np.array_equal(G.col, g["karate/col"])
    assert np.array_equal(G.edges(), golden["karate"]["edges"])
    assert G.number_of_nodes() == 34 and G.number_of_edges() == 78
    deg = G.degree()
    assert [deg[i] for i in sorted(deg)] == golden["karate"]["degrees"].tolist()
    assert gu.file_seed(random.Random(9999999999)) == 102045471  # adsc_Karate.py:79 -> graph_utils.py:150


def test_fast_csr_and_sbm_builder():
    import comemb_b200.utils.graph_utils as gu
    G, block = gu.sbm_graph(2000, 10, 20, seed=1)
    assert len(G) == 2000 and G.rowptr[-1] == G.col.size
    src = np.repeat(np.arange(2000), np.diff(G.rowptr))
    assert (src != G.col).all()                                   # no self loops
    key = src.astype(np.int64) * 2000 + G.col
    assert np.unique(key).size == key.size                        # no duplicates
    assert set(zip(src.tolist(), G.col.tolist())) == set(zip(G.col.tolist(), src.tolist()))  # symmetric
    intra = (block[src] == block[G.col]).mean()
    assert 0.7 < intra < 0.9


def test_io_utils_roundtrip(tmp_path):
    from comemb_b200.utils.IO_utils import load_embedding, load_ground_true, save_embedding
    emb = np.random.RandomState(0).rand(5, 3).astype(np.float32)
    save_embedding(emb, "e", path=str(tmp_path))
    first = open(os.path.join(str(tmp_path), "e.txt")).readline()
    assert first.startswith("1\t") and len(first.split("\t")[1].split(" ")) == 3
    assert np.allclose(load_embedding("e", path=str(tmp_path)), emb)
    labels, k = load_ground_true(os.path.join(ROOT, "tests", "golden"), "karate_zachary")
    assert len(labels) == 34 and k == 2


def test_pairs_per_walk_formula():
    import bench
    assert bench.pairs_of_len(80, 10) == 1490
    assert bench.pairs_of_len(1, 10) == 0 and bench.pairs_of_len(3, 2) == 6
    assert bench.B_PAIR == 7168


def test_oracle_lcg_skip_matches_closed_form():
    # the device skip-ahead composes affine maps; the same algebra in python ints against the oracle's stepping
    a, c, mask = 25214903917, 11, (1 << 48) - 1
    for x0, n in ((123456789, 0), (1, 1), (987654321012, 7450), (2 ** 47 + 5, 12345)):
        A, C, aa, cc, k = 1, 0, a, c, n
        while k:
            if k & 1:
                A, C = (A * aa) & mask, (C * aa + cc) & mask
            cc, aa = ((aa + 1) * cc) & mask, (aa * aa) & mask
            k >>= 1
        assert (A * x0 + C) & mask == O.lcg_advance(x0, n)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU (the reference's compiled train_o2, or the oracle port, on the
    host cores) and prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "sgns_pair_updates_per_sec" and line["value"] > 1e4
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def test_paths_to_rows_rectangular_fast_paths_equal_the_per_path_loop():
    """Edge lists / fixed-length walks are mapped in one vectorised pass (dense ids 1..N without a search, sparse ids
    with searchsorted); both must equal the per-path loop that mirrors prepare_sentences (embedding.py:126-136),
    including OOV drops and the order of the down-sampling draws."""
    from comemb_b200.utils.embedding import paths_to_rows

    class FakeModel(object):
        def __init__(self, ids, probs):
            self.ids, self.probs = np.asarray(ids, np.int64), probs

        def id_index(self):
            return self.ids, np.arange(self.ids.size, dtype=np.int64), self.probs

    rs = np.random.RandomState(0)
    for ids in (np.arange(1, 101), np.array(sorted(rs.choice(500, 100, replace=False)))):
        a = rs.randint(-3, 520, size=(50, 7))
        for probs in (np.ones(ids.size), np.where(np.arange(ids.size) % 3 == 0, 1.0, rs.uniform(0.3, 1.0, ids.size))):
            m = FakeModel(ids, probs)
            np.random.seed(5)
            f1, o1 = paths_to_rows(m, a)
            np.random.seed(5)
            f2, o2 = paths_to_rows(m, [list(r) for r in a])
            assert np.array_equal(f1, f2) and np.array_equal(o1, o2)


def test_community2vec_row_lookup_matches_the_dict_loop():
    """Community2Vec.train maps node ids to rows (community_embeddings.py:63): vectorised for integer ids, dict loop for
    everything else; an unknown id raises KeyError like the reference's `model.vocab[x]`."""
    from comemb_b200.ADSCModel.community_embeddings import _rows_of

    class V(object):
        def __init__(self, i):
            self.index = i

    class WithIndex(object):
        def __init__(self):
            self.vocab = {k: V(i) for i, k in enumerate([5, 3, 9, 1])}

        def id_index(self):
            ids = np.array(sorted(self.vocab))
            return ids, np.array([self.vocab[k].index for k in ids]), None

    class Plain(object):
        def __init__(self):
            self.vocab = {k: V(i) for i, k in enumerate([5, 3, 9, 1])}

    for m in (WithIndex(), Plain()):
        assert _rows_of(m, [9, 1, 5]).tolist() == [2, 3, 0]
        assert _rows_of(m, np.array([3, 3])).tolist() == [1, 1]
        assert _rows_of(m, iter([1, 9])).tolist() == [3, 2]
        with pytest.raises(KeyError):
            _rows_of(m, [9, 2])


def test_graph_rows_are_remapped_to_table_rows():
    """CSR rows follow the first appearance of an id in the edge file, table rows follow the sorted ids
    (model.py:60-65): for the reference's karate file the two differ, and walks from the device walker must be mapped
    before they index the tables (round-1 advisor finding)."""
    import torch
    from comemb_b200.ADSCModel.model import Model
    from comemb_b200.utils import graph_utils as gu
    from comemb_b200.utils.embedding import Vocab
    G = gu.load_adjacencylist(os.path.join(ROOT, "tests", "golden", "karate.adjlist"))
    assert not np.array_equal(G.ids, np.sort(G.ids))  # first-appearance order is not sorted for this file
    m = Model.__new__(Model)
    m.device, m._id_index, m.vocab = "cpu", None, {}
    for row, node in enumerate(sorted(int(i) for i in G.ids if i != 12)):  # node 12 left out of the vocabulary
        m.vocab[node] = Vocab(count=1, index=row, sample_probability=1.0)
    lut = m.graph_row_lut(G)
    assert lut is not None and lut.shape[0] == len(G)
    for csr_row, node in enumerate(G.ids):
        assert int(lut[csr_row]) == (m.vocab[int(node)].index if int(node) in m.vocab else -1)
    walks = torch.tensor([[0, 5, 33, -1], [int(np.flatnonzero(G.ids == 12)[0]), 1, -1, -1]], dtype=torch.int32)
    rows = m.walks_to_rows(G, walks)
    assert rows[0, 3] == -1 and rows[1, 0] == -1 and rows[1, 2] == -1  # padding kept, OOV node dropped
    assert int(rows[0, 1]) == m.vocab[int(G.ids[5])].index
    # sorted dense ids (every synthetic generator): no LUT, tokens pass through untouched
    G2 = gu.from_edge_array_fast(np.array([[1, 2], [2, 3], [3, 1]]), 3)
    m2 = Model.__new__(Model)
    m2.device, m2._id_index = "cpu", None
    m2.vocab = {i: Vocab(count=2, index=i - 1, sample_probability=1.0) for i in (1, 2, 3)}
    assert m2.graph_row_lut(G2) is None and m2.walks_to_rows(G2, walks) is walks


def test_node2vec_loss_matches_the_reference_value():
    """Node2Vec.loss against values produced by the REFERENCE's own Node2Vec.loss (node_embeddings.py:26-31) on its own
    karate tables (tests/golden/make_golden_losses.py -> golden_losses.json): initial table, after the pre-training o1
    epoch, after the first full iteration.  The reference sums in float32, we sum in float64: 1e-5 relative."""
    import json
    import torch
    from comemb_b200.ADSCModel.model import Model
    from comemb_b200.ADSCModel.node_embeddings import Node2Vec
    from comemb_b200.utils.embedding import Vocab
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_losses.json")))["node2vec_loss"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_karate.npz"))
    m = Model.__new__(Model)
    m.device, m._id_index = "cpu", None
    m.vocab = {i + 1: Vocab(count=int(c), index=i, sample_probability=1.0) for i, c in enumerate(g["degrees"])}
    for stage, ref in want.items():
        m.node_embedding = torch.from_numpy(g[stage])
        got = Node2Vec(workers=1, negative=4, lr=0.1).loss(m, g["edges"])
        assert abs(got - ref) <= 1e-5 * abs(ref), (stage, got, ref)
    assert want["init_node"] > want["pre_o1_node"] > want["it0_o2_node"]  # the reference's own training lowers it


def test_nmi_on_tensors_matches_sklearn():
    import torch
    from sklearn.metrics import normalized_mutual_info_score
    from comemb_b200 import evaluation
    rs = np.random.RandomState(0)
    for ka, kb, n in ((2, 2, 34), (5, 7, 1000), (50, 50, 20000), (1, 3, 10)):
        a, b = rs.randint(0, ka, n), rs.randint(0, kb, n)
        b[: n // 2] = a[: n // 2] % kb  # correlated halves
        assert abs(evaluation.nmi(a, torch.as_tensor(b)) - normalized_mutual_info_score(a, b)) < 1e-9
    assert evaluation.nmi(np.arange(10) % 2, torch.as_tensor((np.arange(10) % 2) * 5 + 1)) == pytest.approx(1.0)
    x = np.concatenate([rs.normal(size=(200, 8)) + 6 * rs.normal(size=(1, 8)) for _ in range(4)])
    lab = np.repeat(np.arange(4), 200)
    assert evaluation.community_nmi(torch.as_tensor(x, dtype=torch.float32), lab, method="device") > 0.9


def test_learner_mode_rule():
    """workers=1 replays the reference's order up to ORDERED_AUTO_MAX_UPDATES pair updates and trains lock-free above;
    an explicit mode always wins; workers>1 is lock-free (utils/training_sdg_inner.select_mode)."""
    from comemb_b200.utils import training_sdg_inner as K
    from comemb_b200.ADSCModel.context_embeddings import Context2Vec
    from comemb_b200.ADSCModel.node_embeddings import Node2Vec
    big = K.ORDERED_AUTO_MAX_UPDATES
    assert K.select_mode(None, 1, 10) == K.MODE_ORDERED
    assert K.select_mode(None, 1, big) == K.MODE_ORDERED
    assert K.select_mode(None, 1, big + 1) == K.MODE_HOGWILD
    assert K.select_mode("ordered", 1, 100 * big) == K.MODE_ORDERED
    assert K.select_mode("ordered", 8, 100 * big) == K.MODE_ORDERED
    assert K.select_mode("hogwild", 1, 1) == K.MODE_HOGWILD
    assert K.select_mode(None, 4, 10) == K.MODE_HOGWILD
    assert K.select_mode(K.MODE_ORDERED, 4, 10) == K.MODE_ORDERED
    c = Context2Vec(window_size=5, workers=1)
    assert c._mode(1000) == K.MODE_ORDERED and c._mode(big) == K.MODE_HOGWILD
    assert Context2Vec(workers=1, mode="ordered")._mode(big) == K.MODE_ORDERED
    n = Node2Vec(workers=1)
    assert n._mode(1000) == K.MODE_ORDERED and n._mode(big) == K.MODE_HOGWILD


def test_micro_f1_device_method_matches_sklearn():
    """evaluation.node_classification_micro_f1(method="device") -- multinomial logistic regression by L-BFGS in torch --
    against the sklearn path on the same split: overlapping Gaussian blobs (so that the score is not trivially 1)."""
    import torch
    from comemb_b200.evaluation import node_classification_micro_f1
    rs = np.random.RandomState(3)
    k, n, d = 5, 1500, 16
    centres = rs.randn(k, d) * 0.9
    y = rs.randint(0, k, size=n)
    x = (centres[y] + rs.randn(n, d)).astype(np.float32)
    a = node_classification_micro_f1(x, y, seed=1)
    b = node_classification_micro_f1(torch.from_numpy(x), y, seed=1, method="device")
    assert 0.5 < a < 0.999
    assert abs(a - b) <= 0.01, (a, b)


def test_sgns_objective_evaluator_cpu():
    """evaluation.sgns_objective: (1 + neg) ln 2 (minus the skipped draws) for a zero context table; the positive part equals a direct numpy
    evaluation over the window pairs (None tokens skipped); the pair count matches pyx:494-507's loop."""
    import torch
    from comemb_b200.evaluation import sgns_objective
    rs = np.random.RandomState(5)
    n, d, L, W, neg = 50, 16, 12, 3, 4
    node = torch.from_numpy(rs.normal(size=(n, d)).astype(np.float32) * 0.3)
    ctx = torch.from_numpy(rs.normal(size=(n, d)).astype(np.float32) * 0.3)
    walks = rs.randint(0, n, size=(9, L)).astype(np.int64)
    walks[2, 5:] = 0xFFFFFFFF
    walks[4, 3] = 0xFFFFFFFF
    table = torch.from_numpy(rs.randint(0, n, size=1000).astype(np.int32))
    w32 = torch.from_numpy(walks.astype(np.uint32).view(np.int32))
    obj0, pos0, cnt = sgns_objective(node, torch.zeros_like(ctx), w32, W, table, neg)
    # draws equal to the centre are skipped (2 % of them on a 50-row table): slightly below (1 + neg) ln 2
    assert (1 + 0.9 * neg) * np.log(2) < obj0 <= (1 + neg) * np.log(2) + 1e-9 and abs(pos0 - np.log(2)) < 1e-9
    want, pairs = 0.0, 0
    for path in walks:
        for i in range(L):
            if path[i] == 0xFFFFFFFF:
                continue
            for j in range(max(0, i - W), min(L, i + W + 1)):
                if j == i or path[j] == 0xFFFFFFFF:
                    continue
                f = float(node[path[j]].double() @ ctx[path[i]].double())
                want += np.log1p(np.exp(-f))
                pairs += 1
    obj, pos, cnt = sgns_objective(node, ctx, w32, W, table, neg)
    assert cnt == pairs and abs(pos - want / pairs) < 1e-9
    assert obj > pos


def test_c_abi_argument_errors_return_status_codes_without_a_gpu():
    """The argument checks of the C ABI run before any CUDA call: null pointers / negative sizes give COMEMB_E_ARG (-1),
    shapes no kernel covers COMEMB_E_UNSUPPORTED (-2), a launch before comemb_init() COMEMB_E_NOINIT (-3), each with a
    message -- the error behaviour the host mirror turns into exceptions (ComembError)."""
    from comemb_b200 import _lib
    lib = _lib.load()
    assert lib.comemb_gmm_mstep(None, 10, 128, None, None, 3, None, None) == -1
    assert lib.comemb_gmm_mstep(1, 10, 64, 1, 1, 3, 1, None) == -2
    assert lib.comemb_gmm_estep(None, 10, 128, None, None, 3, None, None) == -1
    assert lib.comemb_gmm_estep(1, 10, 64, 1, 1, 3, 1, None) == -2
    assert lib.comemb_set_max_warps(-1) == -1 and lib.comemb_set_tuning(-1, 0, 0) == -1
    st = lib.comemb_o2_walks(None, None, 10, 128, None, None, 1, None, 0, None, 0, None, 0, 5, 5, 0.1, 1.0, 0, 0, None, None)
    assert st in (-1, -3)
    for code, word in ((-1, b"invalid argument"), (-2, b"unsupported"), (-3, b"comemb_init")):
        assert word in lib.comemb_error_string(code)
    with pytest.raises(_lib.ComembError):
        _lib.check(-2)
