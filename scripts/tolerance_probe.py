"""Measured values behind the stated Hogwild tolerances of tests/test_gpu_parity.py (run on a B200; prints one line per
quantity so that the tolerances in the tests can be set at a small multiple of what is observed)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import comemb_b200.utils.training_sdg_inner as K  # noqa: E402
from oracle import oracle as O  # noqa: E402

K.init()
O.build()


def dev(a):
    a = np.ascontiguousarray(a)
    a = a.view(np.int32) if a.dtype == np.uint32 else (a.view(np.int64) if a.dtype == np.uint64 else a)
    return torch.from_numpy(a).cuda()


# (1) single-warp Hogwild with red.add vs the oracle in warp order: max |diff| per golden case
for name in ["o2_d128_small", "o2_d2_karate_default", "o2_d100_tail", "o2_d160_blk32", "o2_d64_none_ragged", "o2_d256"]:
    c = cases.O2_CASES[name]
    node, ctx, table, walks = cases.o2_inputs(c)
    seeds = O.seeds_from_numpy(np.random.RandomState(3), len(walks))
    dn, dc, dt = dev(node), dev(ctx), dev(table)
    for w, s in zip(walks, seeds):
        if len(w):
            K.o2_batch(dn, dc, dev(w), dev(np.array([0, len(w)], np.int64)), dev(np.array([s], np.uint64)), c["lr"],
                       c["neg"], c["W"], dt, alpha=c["lam"], mode=K.MODE_HOGWILD, flags=K.F_ATOMIC)
    flat, off = cases.flatten_walks(walks)
    O.o2_walks(node, ctx, flat, off, seeds, c["lr"], c["neg"], c["W"], table, c["lam"], O.DOT_WARP)
    print("atomic single warp %-22s max|d node| %.3e  max|d ctx| %.3e  (table scale %.2f)" % (
        name, np.abs(dn.cpu().numpy() - node).max(), np.abs(dc.cpu().numpy() - ctx).max(), np.abs(node).max()), flush=True)

# (2) ORDERED vs HOGWILD on the small SBM of test_hogwild_training_quality_matches_ordered_on_sbm, several seeds
import comemb_b200.utils.graph_utils as gu  # noqa: E402
from comemb_b200 import evaluation  # noqa: E402
n, k, d, L, W = 2000, 5, 128, 30, 5
G, block = gu.sbm_graph(n, k, 20, p_in=0.9, seed=11)
walks, lens = gu.build_deepwalk_corpus(G, 2, L, alpha=0.0, seed=3, mode=gu.MODE_HOGWILD, return_device=True)
nw = walks.shape[0]
off = torch.arange(nw + 1, dtype=torch.int64, device="cuda") * L
table = dev(O.make_table(np.diff(G.rowptr).astype(np.float64), 100000))
for seed in (5, 6, 7):
    node0 = (np.random.RandomState(seed).uniform(-1, 1, (n, d)) * 0.18).astype(np.float32)
    seeds = dev(O.seeds_from_numpy(np.random.RandomState(seed), nw))
    res = {}
    for tag, mode, flags in (("ordered", K.MODE_ORDERED, 0), ("plain", K.MODE_HOGWILD, 0), ("atomic", K.MODE_HOGWILD, K.F_ATOMIC)):
        a, b = dev(node0), torch.zeros((n, d), device="cuda")
        for epoch in range(2):
            K.o2_batch(a, b, walks.reshape(-1), off, seeds, 0.05, 5, W, table, mode=mode, flags=flags)
        loss, pairs = K.o2_pos_loss(a, b, walks.reshape(-1), off, W)
        res[tag] = (loss / pairs, evaluation.community_nmi(a, block, k=k, method="device"))
    print("sbm2000 seed %d: " % seed + "  ".join("%s loss %.4f nmi %.3f" % (t, *res[t]) for t in res) +
          "   rel loss diff atomic %.3f plain %.3f" % (abs(res["atomic"][0] - res["ordered"][0]) / res["ordered"][0],
                                                     abs(res["plain"][0] - res["ordered"][0]) / res["ordered"][0]), flush=True)
