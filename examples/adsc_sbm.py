#!/usr/bin/env python3
"""The reference's outer loop (adsc_Karate.py:105-138: o1 epoch, o2 epoch, GMM fit, 5 x o3) at BASELINE configs[1]
scale -- synthetic SBM, 100K nodes / 2M edges / 50 communities, d=128 -- entirely on one B200: CSR walker, Hogwild
o1/o2 kernels, device GMM, o3 kernel.  Prints per-stage wall-clock and the community NMI of the GMM assignment.

    python examples/adsc_sbm.py [--n 100000] [--blocks 50] [--walks 10]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from comemb_b200.ADSCModel.model import Model  # noqa: E402
from comemb_b200.ADSCModel.context_embeddings import Context2Vec  # noqa: E402
from comemb_b200.ADSCModel.node_embeddings import Node2Vec  # noqa: E402
from comemb_b200.ADSCModel.community_embeddings import Community2Vec  # noqa: E402
import comemb_b200.utils.graph_utils as gu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--blocks", type=int, default=50)
    ap.add_argument("--walks", type=int, default=10)
    ap.add_argument("--size", type=int, default=128)
    args = ap.parse_args()
    t = time.perf_counter()
    G, block = gu.sbm_graph(args.n, args.blocks, 40, seed=12345)
    np.random.seed(1)
    model = Model(G.degree(), size=args.size, table_size=5000000, k=args.blocks)
    model.node_embedding.mul_(0.05)  # the reference's U(-1,1) init saturates sigma at d=128
    print("graph + model: %.1fs (%d nodes, %d edges)" % (time.perf_counter() - t, len(G), G.number_of_edges()))
    workers = 64  # > 1 -> HOGWILD
    n2v = Node2Vec(workers=workers, negative=5, lr=0.025)
    c2v = Context2Vec(window_size=10, workers=workers, negative=5, lr=0.025)
    com = Community2Vec(model, lr=0.025, reg_covar=1e-4, gmm_backend="device")
    edges = G.edges()

    def stage(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        print("%-28s %.3fs" % (name, time.perf_counter() - t0))

    def o2_epoch():
        walks, lens = gu.build_deepwalk_corpus(G, args.walks, 80, alpha=0, seed=7, mode=gu.MODE_HOGWILD,
                                               return_device=True)
        c2v.train(model, paths=(walks, lens), total_nodes=walks.numel(), alpha=1.0)

    stage("o1 epoch (2M edges)", lambda: n2v.train(model, edges=edges, iter=1))
    stage("o2 epoch (%d walks/node)" % args.walks, o2_epoch)
    stage("GMM fit (device, n_init=10)", lambda: com.fit(model))
    stage("o3 x5", lambda: com.train(G.nodes(), model, beta=0.1, iter=5))
    from sklearn.metrics import normalized_mutual_info_score as nmi
    pred = model.pi.argmax(1).cpu().numpy()
    print("community NMI (GMM assignment vs SBM blocks): %.3f" % nmi(block, pred))


if __name__ == "__main__":
    main()
