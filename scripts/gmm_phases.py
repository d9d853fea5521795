"""Where the device GMM fit spends its time (N1): k-means init, E-step, M-step (means/covariances), precision factors.

    python scripts/gmm_phases.py [--n 100000] [--k 50] [--separation 0.35]
Clustered points like a trained SBM node table (K blobs).  Times are CUDA-event milliseconds summed over one fit."""
import argparse
import json
import os
import sys
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__  # noqa: E402,F401
from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture  # noqa: E402


class Timed(DeviceGaussianMixture):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.ms = defaultdict(float)
        self.calls = defaultdict(int)

    def _t(self, name, fn, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a)
        e1.record()
        torch.cuda.synchronize()
        self.ms[name] += e0.elapsed_time(e1)
        self.calls[name] += 1
        return r

    def _estimate_parameters(self, X, resp):
        return self._t("m_step_parameters", super()._estimate_parameters, X, resp)

    def _log_prob_resp(self, X):
        return self._t("e_step", super()._log_prob_resp, X)

    def _kmeans_resp(self, X, gen):
        return self._t("kmeans_init", super()._kmeans_resp, X, gen)

    def _precision_cholesky(self, covs):
        return self._t("precision_cholesky", DeviceGaussianMixture._precision_cholesky, covs)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--separation", type=float, default=0.35)
    ap.add_argument("--n-init", type=int, default=1)
    a = ap.parse_args()
    rs = np.random.RandomState(0)
    centres = rs.randn(a.k, a.d).astype(np.float32) * a.separation
    lab = rs.randint(0, a.k, size=a.n)
    x = torch.from_numpy(centres[lab] + rs.randn(a.n, a.d).astype(np.float32) * 0.25).cuda()
    for sparse in (True, False):
        gm = Timed(n_components=a.k, reg_covar=1e-6, n_init=a.n_init, sparse_m_step=sparse)
        gm.fit(x)  # warm-up (allocator, cuSOLVER handles)
        gm.ms.clear(); gm.calls.clear()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); gm.fit(x); t1.record(); torch.cuda.synchronize()
        print(json.dumps({"n": a.n, "k": a.k, "sparse_m_step": sparse, "fit_ms": round(t0.elapsed_time(t1), 2),
                          "n_iter": gm.n_iter_, "phases_ms": {k: round(v, 2) for k, v in gm.ms.items()},
                          "calls": dict(gm.calls)}), flush=True)
