import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from oracle import oracle as O
import comemb_b200.utils.training_sdg_inner as K
from test_gpu_parity import _fast_sg_inputs, dev, host
K.init()
W, neg, lr, l1, lam2 = 5, 5, 0.025, 0.9, 0.3
node, ctx, table, mu, inv, pi, walks = _fast_sg_inputs(7)
node0 = node.copy()
rs = np.random.RandomState(8)
seeds = O.seeds_from_numpy(rs, len(walks))
dn, dc, dt = dev(node), dev(ctx), dev(table)
dmu, dinv, dpi = dev(mu), dev(inv), dev(pi)
for wi_, (w, s) in enumerate(zip(walks, seeds)):
    before = host(dn).copy()
    K.sg_batch(dn, dc, dev(w), dev(np.array([0, len(w)], np.int64)), None, dev(np.array([s], np.uint64)), lr, neg, W, dt, dmu, dinv, dpi, l1, lam2, 0, mode=K.MODE_HOGWILD)
    O.train_sg(node, ctx, np.ascontiguousarray(w), None, lr, neg, W, table, mu, inv, pi, l1, lam2, 0, int(s), O.DOT_WARP)
    diff = np.abs(host(dn) - node)
    rows = np.flatnonzero(diff.max(1) > 1e-4)
    print("walk", wi_, "max", diff.max(), "mean", diff.mean(), "rows>1e-4:", len(rows), "frac coords", (diff > 1e-4).mean())
    if wi_ == 0:
        for r in rows[:6]:
            pos = np.flatnonzero(w == r)
            c = np.flatnonzero(pi[r])
            print("  row", r, "walk pos", pos, "comm", c, "n coords off", int((diff[r] > 1e-4).sum()), "max", diff[r].max(), "moved", np.abs(node[r]-node0[r]).max())
        break
