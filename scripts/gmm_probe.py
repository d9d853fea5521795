import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__
from comemb_b200.ADSCModel import gmm_device as G
rs = np.random.RandomState(0)
k, d, n = 50, 128, 100000
centres = rs.randn(k, d).astype(np.float32) * 0.05
lab = rs.randint(0, k, size=n)
x = torch.from_numpy(centres[lab] + rs.randn(n, d).astype(np.float32) * 0.25).cuda()
gm = G.DeviceGaussianMixture(n_components=k, n_init=1, max_iter=20)
gm.fit(x)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.time() - t0
gen = torch.Generator(device=x.device); gen.manual_seed(0)
t0 = time.time(); resp = gm._kmeans_resp(x, gen); tick("kmeans", t0)
t0 = time.time(); nk, gm.means_, gm.covariances_ = gm._estimate_parameters(x, resp); tick("params", t0)
gm.weights_ = nk / n
t0 = time.time(); gm.precisions_cholesky_ = gm._precision_cholesky(gm.covariances_); tick("chol", t0)
for it in range(20):
    t0 = time.time(); norm, log_resp = gm._log_prob_resp(x); tick("estep", t0)
    t0 = time.time(); r = G.torch_exp(log_resp); tick("exp", t0)
    t0 = time.time(); nk, gm.means_, gm.covariances_ = gm._estimate_parameters(x, r); tick("params", t0)
    t0 = time.time(); gm.weights_ = nk / nk.sum(); tick("weights", t0)
    t0 = time.time(); gm.precisions_cholesky_ = gm._precision_cholesky(gm.covariances_); tick("chol", t0)
    t0 = time.time(); lower = float(norm.mean()); tick("lower", t0)
print(json.dumps({a: round(b * 1e3, 1) for a, b in T.items()}))
# inside params: sparse attempt cost
t0 = time.time(); nz = r != 0; per_k = nz.sum(0); m = int(per_k.max()); tick("sparse_probe", t0)
print("m", m, json.dumps({a: round(b * 1e3, 1) for a, b in T.items() if a == "sparse_probe"}))
t0 = time.time(); gm2 = G.DeviceGaussianMixture(n_components=k, n_init=1, max_iter=20); gm2.fit(x); tick("fit20", t0)
print("fit20 ms", round(T["fit20"] * 1e3, 1), gm2.n_iter_)
import cProfile, pstats
gm3 = G.DeviceGaussianMixture(n_components=k, n_init=1, max_iter=20)
pr = cProfile.Profile(); pr.enable(); gm3.fit(x); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
