#!/usr/bin/env python3
"""The reference's driver (adsc_Karate.py:34-149) on the B200 hot path: same hyper-parameters, same call order
(pre-train o1 -> o2, then one outer iteration o1 -> o2 -> GMM fit -> 5 x o3), only the import roots change
(INTEGRATION.md option B).  Plotting is dropped; the community NMI against the Zachary labels is printed instead.

    python examples/adsc_karate.py [--size 128] [--workers 1]      # workers=1: the reference's results bit for bit
"""
import argparse
import os
import random
import sys
import timeit

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from comemb_b200.ADSCModel.model import Model  # noqa: E402
from comemb_b200.ADSCModel.context_embeddings import Context2Vec  # noqa: E402
from comemb_b200.ADSCModel.node_embeddings import Node2Vec  # noqa: E402
from comemb_b200.ADSCModel.community_embeddings import Community2Vec  # noqa: E402
import comemb_b200.utils.graph_utils as graph_utils  # noqa: E402
import comemb_b200.utils.IO_utils as io_utils  # noqa: E402
from comemb_b200.evaluation import community_nmi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)  # the reference hard-codes 2 (adsc_Karate.py:39)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--out", default=None, help="directory for the embedding text files (IO_utils.save_embedding)")
    args = ap.parse_args()
    data = os.path.join(ROOT, "tests", "golden")
    number_walks, walk_length, window_size, negative = 10, 20, 3, 4      # adsc_Karate.py:37-47
    alpha, beta, lr, reg_covar = 1.0, 0.01, 0.1, 0.00001
    np.random.seed(1)

    G = graph_utils.load_adjacencylist(os.path.join(data, "karate.adjlist"), True)           # :57
    model = Model(G.degree(), size=args.size, table_size=5000000, input_file="karate_zachary", path_labels=data)
    mode = graph_utils.MODE_ORDERED if args.workers == 1 else graph_utils.MODE_HOGWILD
    walks, lens = graph_utils.build_deepwalk_corpus(G, number_walks, walk_length, alpha=0,     # :75-80
                                                    seed=graph_utils.file_seed(random.Random(9999999999)), mode=mode)
    paths = [G.ids[w[:n].astype(np.int64)] for w, n in zip(walks, lens)]

    node_learner = Node2Vec(workers=args.workers, negative=negative, lr=lr)
    cont_learner = Context2Vec(window_size=window_size, workers=args.workers, negative=negative, lr=lr)
    com_learner = Community2Vec(model, reg_covar=reg_covar, lr=lr)
    context_total_path = G.number_of_nodes() * number_walks * walk_length
    edges = np.array(G.edges())

    node_learner.train(model, edges=edges, iter=1, chunksize=20)                               # :105-113 pre-training
    cont_learner.train(model, paths=paths, total_nodes=context_total_path, alpha=alpha, chunksize=20)
    start = timeit.default_timer()
    node_learner.train(model, edges=edges, iter=1, chunksize=20)                               # :125-137
    cont_learner.train(model, paths=paths, total_nodes=context_total_path, alpha=alpha, chunksize=20)
    com_learner.fit(model)
    com_learner.train(G.nodes(), model, beta, chunksize=20, iter=5)
    print("outer iteration: %.2fs" % (timeit.default_timer() - start))
    if args.out:
        io_utils.save_embedding(model.node_embedding, "karate_alpha-%s_beta-%s" % (alpha, beta), path=args.out)
    x = model.node_embedding.cpu().numpy()
    print("o1 loss %.3f  NMI vs karate_zachary.labels: %.3f" % (node_learner.loss(model, edges),
                                                               community_nmi(x, model.ground_true, k=model.k)))


if __name__ == "__main__":
    main()
