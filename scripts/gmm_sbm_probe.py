"""How sparse are the responsibilities during the GMM fit on real (one-epoch SBM) embeddings?"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comemb_b200.ADSCModel.model import Model  # noqa: E402
from comemb_b200.ADSCModel.context_embeddings import Context2Vec  # noqa: E402
from comemb_b200.ADSCModel.node_embeddings import Node2Vec  # noqa: E402
from comemb_b200.ADSCModel import gmm_device  # noqa: E402
import comemb_b200.utils.graph_utils as gu  # noqa: E402

G, block = gu.sbm_graph(100000, 50, 40, seed=12345)
np.random.seed(1)
model = Model(G.degree(), size=128, table_size=5000000, k=50)
model.node_embedding.mul_(0.05)
Node2Vec(workers=64, negative=5, lr=0.025).train(model, edges=G.edges(), iter=1)
walks, lens = gu.build_deepwalk_corpus(G, 10, 80, alpha=0, seed=7, mode=gu.MODE_HOGWILD, return_device=True)
Context2Vec(window_size=10, workers=64, negative=5, lr=0.025).train(model, paths=(walks, lens), total_nodes=walks.numel(),
                                                                   alpha=1.0)
X = model.node_embedding.detach()
stats = []
orig = gmm_device.DeviceGaussianMixture._covariances_sparse


def probe(self, X_, resp, nk, means):
    out = orig(self, X_, resp, nk, means)
    nz = resp != 0
    stats.append((float(nz.float().mean()), int(nz.sum(0).max()), out is not None))
    return out


gmm_device.DeviceGaussianMixture._covariances_sparse = probe
times = {}


def timed(name):
    f = getattr(gmm_device.DeviceGaussianMixture, name)

    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = f(self, *a, **k)
        torch.cuda.synchronize()
        times.setdefault(name, []).append(1e3 * (time.perf_counter() - t0))
        return out
    setattr(gmm_device.DeviceGaussianMixture, name, g)


for nm in ("_log_prob_resp", "_estimate_parameters", "_kmeans_resp"):
    timed(nm)
gmm_device.DeviceGaussianMixture._precision_cholesky = staticmethod(
    (lambda f: (lambda covs: (torch.cuda.synchronize(), time.perf_counter(), f(covs), torch.cuda.synchronize(),
                              times.setdefault("_precision_cholesky", []).append(0.0))[2]))(
        gmm_device.DeviceGaussianMixture._precision_cholesky))
for reg in (1e-4,):
    torch.cuda.synchronize(); t = time.perf_counter()
    gm = gmm_device.DeviceGaussianMixture(n_components=50, reg_covar=reg, n_init=2, random_state=0).fit(X)
    torch.cuda.synchronize()
    print("fit n_init=2: %.2f s, iterations of the best init %d, converged %s" % (time.perf_counter() - t, gm.n_iter_, gm.converged_))
print("M-steps: %d, sparse path taken in %d" % (len(stats), sum(s[2] for s in stats)))
for s in stats[:6] + stats[-3:]:
    print("  non-zero fraction %.4f  max per component %d  sparse %s" % s)
for k, v in times.items():
    print("%-22s calls %3d  mean %.2f ms  max %.2f ms" % (k, len(v), sum(v) / len(v), max(v)))
