#!/usr/bin/env python3
"""Generate the golden vectors by running the REFERENCE ITSELF (authoring container only).

  * o1/o2: the reference's compiled Cython module (oracle/_ref/tuned, built from
    /root/reference/utils/training_sdg_inner.pyx by oracle/build_ref.py) -- train_o1 (pyx:407), train_o2 (pyx:454).
  * o3: /root/reference/ADSCModel/community_embeddings.py Community2Vec.train (:61-77), imported in place.
  * table: /root/reference/ADSCModel/model.py Model.make_table (:97-122).
  * walks: /root/reference/utils/graph_utils.py build_deepwalk_corpus_iter (:191-197) on data/karate, with the
    networkx>=2 shim (`neighbors` returns a list) described in SURVEY.md section 8c.
  * karate: config 1 -- Model + Node2Vec.train + Context2Vec.train + Community2Vec.train at d=128 through the
    reference's own learner classes (workers=1), GMM parameters recorded as inputs.

Outputs (committed): tests/golden/golden_sgd.npz, golden_sg.npz (legacy fused train_sg, from the stale Cython
kernel rebuilt by oracle/build_ref_legacy.py), golden_walks.npz, golden_karate.npz, golden_meta.json.
Run:  python tests/golden/make_golden.py        (needs /root/reference; never runs on the GPU box)
"""
import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import build_ref  # noqa: E402


def to_path(tokens):
    return [None if int(t) == cases.TOKEN_NONE else O.RefVocab(int(t)) for t in tokens]


def gen_sgd(ref, out):
    for name, c in cases.O2_CASES.items():
        node, ctx, table, walks = cases.o2_inputs(c)
        np.random.seed(c["seed"] + 7)  # the stream pyx:477 draws the per-call seeds from
        work = np.zeros(c["d"], np.float32)
        tot = 0
        for w in walks:
            tot += ref.train_o2(node, ctx, to_path(w), c["lr"], c["neg"], c["W"], table, py_alpha=c["lam"],
                                py_size=c["d"], py_work=work)
        out[name + "/node"] = node
        out[name + "/ctx"] = ctx
        out[name + "/ret"] = np.int64(tot)
    for name, c in cases.O1_CASES.items():
        node, table, edges = cases.o1_inputs(c)
        np.random.seed(c["seed"] + 7)
        work = np.zeros(c["d"], np.float32)
        tot = 0
        for e in edges:
            tot += ref.train_o1(node, to_path(e), c["lr"], c["neg"], table, py_size=c["d"], py_work=work)
        out[name + "/node"] = node
        out[name + "/ret"] = np.int64(tot)


def gen_sg(out):
    """Legacy fused train_sg: the stale Cython kernel rebuilt by oracle/build_ref_legacy.py."""
    from oracle import build_ref_legacy
    build_ref_legacy.main()
    leg = O.load_ref("legacy")
    assert leg.FAST_VERSION == 0
    for name, c in cases.SG_CASES.items():
        node, ctx, table, mu, inv, pi, walks = cases.sg_inputs(c)
        d = c["d"]
        negemb = node if c["isnode"] else ctx
        np.random.seed(c["seed"] + 7)
        w = [np.zeros(d, np.float32) for _ in range(3)] + [np.zeros(d * d, np.float32)]
        tot = 0
        for path in walks:
            tot += leg.train_sg(node, negemb, to_path(path), c["lr"], c["neg"], c["W"], table, mu, inv, pi, c["K"], inv,
                                py_lambda1=c["l1"], py_lambda2=c["l2"], py_size=d, py_work=w[0], py_work_o3=w[1],
                                py_work1_o3=w[2], py_work2_o3=w[3], py_is_node_embedding=c["isnode"])
        out[name + "/node"] = node
        out[name + "/ctx"] = ctx
        out[name + "/ret"] = np.int64(tot)


def gen_twin(out):
    """The reference's pure-Python fallback train_sg (utils/embedding.py:15-72), imported in place."""
    import utils.embedding as E
    assert E.train_sg.__module__ == "utils.embedding"  # the fallback, not a compiled kernel
    for name, c in cases.TWIN_CASES.items():
        node, ctx, table, mu, inv, pi, walks = cases.sg_inputs(c)
        negemb = node if c["isnode"] else ctx
        np.random.seed(c["seed"] + 7)
        tot = 0
        for path in walks:
            tot += E.train_sg(node, negemb, to_path(path), c["lr"], c["neg"], c["W"], table, mu, inv, pi, c["K"], inv,
                              py_lambda1=c["l1"], py_lambda2=c["l2"], py_size=c["d"],
                              py_is_node_embedding=c["isnode"])
        out[name + "/node"] = node
        out[name + "/ctx"] = ctx
        out[name + "/ret"] = np.int64(tot)


class _FakeModel(object):
    pass


def gen_o3(out):
    from ADSCModel.community_embeddings import Community2Vec
    for name, c in cases.O3_CASES.items():
        node, mu, inv, pi, rows = cases.o3_inputs(c)
        m = _FakeModel()
        m.k = c["K"]
        m.node_embedding = node
        m.centroid, m.inv_covariance_mat, m.pi = mu, inv, pi
        m.vocab = {int(r) + 1: O.RefVocab(int(r)) for r in range(c["N"])}
        learner = Community2Vec(m, lr=c["lr"])
        learner.train([int(r) + 1 for r in rows], m, c["beta"], chunksize=20, iter=c["iters"])
        out[name + "/node"] = m.node_embedding


def rle(table):
    """Monotone table -> (values, first index of each run)."""
    t = np.asarray(table)
    starts = np.flatnonzero(np.concatenate([[True], t[1:] != t[:-1]]))
    return t[starts].astype(np.uint32), starts.astype(np.int64)


def gen_table(out):
    from ADSCModel.model import Model
    rs = np.random.RandomState(400)
    for name, (counts, size) in {
        "table_small": (rs.randint(1, 20, size=50), 10007),
        "table_powerlaw": ((rs.pareto(1.5, size=400) * 3 + 1).astype(np.int64), 250000),
    }.items():
        m = _FakeModel()
        m.vocab = {i + 1: O.RefVocab(i, int(cnt)) for i, cnt in enumerate(counts)}
        m.vocab_size = len(counts)
        m.table_size = size
        Model.make_table(m)
        vals, starts = rle(m.table)
        out[name + "/counts"] = np.asarray(counts, np.int64)
        out[name + "/size"] = np.int64(size)
        out[name + "/vals"] = vals
        out[name + "/starts"] = starts


def load_karate_graph():
    import networkx as nx
    import utils.graph_utils as gu

    class ListNeighborsGraph(nx.Graph):  # networkx>=2 returns iterators; the reference needs len()/indexing
        def neighbors(self, n):
            return list(nx.Graph.neighbors(self, n))

    G = gu.load_adjacencylist(os.path.join(O.HERE, "..", "tests", "golden", "karate.adjlist"), True)
    G.__class__ = ListNeighborsGraph
    return G, gu


def graph_to_csr(G):
    """CSR whose row order is list(G.nodes()) and column order list(G.neighbors(v)) (row numbers, not ids)."""
    nodes = list(G.nodes())
    pos = {v: i for i, v in enumerate(nodes)}
    rowptr = np.zeros(len(nodes) + 1, np.int64)
    col = []
    for i, v in enumerate(nodes):
        nb = list(G.neighbors(v))
        col.extend(pos[u] for u in nb)
        rowptr[i + 1] = len(col)
    return np.asarray(nodes, np.int64), rowptr, np.asarray(col, np.uint32)


def gen_walks(out):
    G, gu = load_karate_graph()
    ids, rowptr, col = graph_to_csr(G)
    out["karate/ids"], out["karate/rowptr"], out["karate/col"] = ids, rowptr, col
    pos = {int(v): i for i, v in enumerate(ids)}
    for name, (num_paths, L, alpha, seed) in {
        "walks_a0": (10, 20, 0.0, 102045471),      # the karate driver's own file seed (adsc_Karate.py:79)
        "walks_a02": (3, 40, 0.2, 12345),
        "walks_len1": (2, 1, 0.0, 7),
        "walks_bigseed": (1, 10, 0.5, 9999999999),
    }.items():
        walks = list(gu.build_deepwalk_corpus_iter(G, num_paths, L, alpha=alpha, rand=random.Random(seed)))
        arr = np.full((len(walks), L), cases.TOKEN_NONE, np.uint32)
        for i, w in enumerate(walks):
            arr[i, :len(w)] = [pos[int(v)] for v in w]
        out[name + "/params"] = np.asarray([num_paths, L, seed], np.int64)
        out[name + "/alpha"] = np.float64(alpha)
        out[name + "/walks"] = arr
    out["file_seed_parent"] = np.int64(9999999999)
    out["file_seed"] = np.int64(random.Random(9999999999).randint(0, 2 ** 31))
    return G, gu


def gen_karate(out, G, gu):
    """Config 1 (BASELINE.json configs[0]) at d=128 through the reference's learner classes, one outer iteration as
    adsc_Karate.py:105-137 orders it: o1 epoch, o2 epoch (pre-training), o1, o2, GMM fit, 5x o3."""
    from ADSCModel.model import Model
    from ADSCModel.node_embeddings import Node2Vec
    from ADSCModel.context_embeddings import Context2Vec
    from ADSCModel.community_embeddings import Community2Vec
    d, number_walks, walk_length, window, negative = 128, 10, 20, 3, 4
    alpha, beta, lr = 1.0, 0.01, 0.1
    np.random.seed(2024)
    model = Model(dict(G.degree()), size=d, table_size=5000000, input_file="karate_zachary", path_labels=HERE)
    out["init_node"] = model.node_embedding.copy()
    vals, starts = rle(model.table)
    out["table_vals"], out["table_starts"] = vals, starts
    out["degrees"] = np.asarray([model.vocab[i].count for i in sorted(model.vocab)], np.int64)
    walks = list(gu.build_deepwalk_corpus_iter(G, number_walks, walk_length, alpha=0,
                                               rand=random.Random(102045471)))
    edges = np.array(G.edges())
    out["edges"] = edges.astype(np.int64)
    out["walks_ids"] = np.asarray(walks, np.int64)
    node_learner = Node2Vec(workers=1, negative=negative, lr=lr)
    cont_learner = Context2Vec(window_size=window, workers=1, negative=negative, lr=lr)
    cont_learner.alpha = alpha  # dodge the reference's `self.alpha` AttributeError (context_embeddings.py:92)
    com_learner = Community2Vec(model, reg_covar=1e-5, lr=lr)
    np.random.seed(77)
    for stage in ("pre", "it0"):
        node_learner.train(model, edges=edges, iter=1, chunksize=20)
        out[stage + "_o1_node"] = model.node_embedding.copy()
        cont_learner.train(model, paths=[np.asarray(w) for w in walks], total_nodes=len(walks) * walk_length,
                           alpha=alpha, chunksize=20)
        out[stage + "_o2_node"] = model.node_embedding.copy()
        out[stage + "_o2_ctx"] = model.context_embedding.copy()
    com_learner.fit(model)
    out["centroid"], out["inv_cov"], out["pi"] = model.centroid, model.inv_covariance_mat, model.pi
    # iter=5 of adsc_Karate.py:137 as five iter=1 calls (the loop body carries no state across iterations,
    # community_embeddings.py:62-77) so every intermediate table is recorded: with reg_covar=1e-5 the inverse
    # covariances reach 1e5 and lr*beta/K*eig >> 2, i.e. the reference's own 5-step result is chaotic in fp32 and
    # only single steps can be compared between two summation orders.
    for it in range(5):
        com_learner.train(list(G.nodes()), model, beta, chunksize=20, iter=1)
        out["o3_iter%d_node" % it] = model.node_embedding.copy()
    out["final_node"] = model.node_embedding.copy()
    out["labels"] = np.asarray(model.ground_true, np.int64)


def main():
    build_ref.main()
    ref = O.load_ref("tuned", with_python_sources=True)
    assert ref.FAST_VERSION == 0
    meta = {"FAST_VERSION": int(ref.FAST_VERSION)}
    import scipy
    import Cython
    import networkx
    import sklearn
    from threadpoolctl import threadpool_info
    meta["versions"] = dict(numpy=np.__version__, scipy=scipy.__version__, cython=Cython.__version__,
                            networkx=networkx.__version__, sklearn=sklearn.__version__,
                            python=sys.version.split()[0])
    meta["blas"] = [dict(prefix=i.get("prefix"), version=i.get("version"), architecture=i.get("architecture"))
                    for i in threadpool_info()]
    sg_out = {}
    gen_sg(sg_out)
    np.savez_compressed(os.path.join(HERE, "golden_sg.npz"), **sg_out)
    ref = O.load_ref("tuned", with_python_sources=True)
    gen_twin(sg_out)
    np.savez_compressed(os.path.join(HERE, "golden_sg.npz"), **sg_out)
    for fname, gen in (("golden_sgd.npz", lambda o: (gen_sgd(ref, o), gen_o3(o), gen_table(o))),):
        out = {}
        gen(out)
        np.savez_compressed(os.path.join(HERE, fname), **out)
    out = {}
    G, gu = gen_walks(out)
    np.savez_compressed(os.path.join(HERE, "golden_walks.npz"), **out)
    out = {}
    gen_karate(out, G, gu)
    np.savez_compressed(os.path.join(HERE, "golden_karate.npz"), **out)
    for f in ("golden_sgd.npz", "golden_sg.npz", "golden_walks.npz", "golden_karate.npz"):
        meta[f] = hashlib.sha256(open(os.path.join(HERE, f), "rb").read()).hexdigest()
    json.dump(meta, open(os.path.join(HERE, "golden_meta.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
