"""Seeded input recipes shared by the golden-vector generator (make_golden.py, runs the compiled reference in the
authoring container) and by the tests (which rebuild the same inputs and compare against the stored outputs).

All randomness comes from legacy `np.random.RandomState(seed)` streams, which numpy keeps bit-stable across versions.
"""
import numpy as np

TOKEN_NONE = 0xFFFFFFFF

# name -> parameters.  `none_every`: every n-th token of a walk is None (pyx:485-486); `ragged`: walk lengths vary.
O2_CASES = {
    "o2_d128_small": dict(d=128, N=50, nw=20, L=40, W=5, neg=5, seed=100, lr=0.025, lam=1.0, scale=0.5),
    "o2_d128_sbm_shape": dict(d=128, N=1000, nw=40, L=80, W=10, neg=5, seed=101, lr=0.025, lam=1.0, scale=0.5),
    "o2_d128_init_like_model": dict(d=128, N=200, nw=30, L=80, W=10, neg=5, seed=102, lr=0.1, lam=1.0, scale=1.0,
                                    ctx_zero=True),
    "o2_d2_karate_default": dict(d=2, N=34, nw=60, L=20, W=3, neg=4, seed=103, lr=0.1, lam=1.0, scale=1.0),
    "o2_d100_tail": dict(d=100, N=60, nw=10, L=30, W=4, neg=3, seed=104, lr=0.05, lam=0.7, scale=0.5),
    "o2_d160_blk32": dict(d=160, N=60, nw=10, L=30, W=4, neg=3, seed=105, lr=0.05, lam=1.0, scale=0.5),
    "o2_d16": dict(d=16, N=60, nw=10, L=30, W=4, neg=3, seed=106, lr=0.05, lam=1.0, scale=1.0),
    "o2_d64_none_ragged": dict(d=64, N=80, nw=25, L=30, W=4, neg=5, seed=107, lr=0.05, lam=1.0, scale=0.5,
                               none_every=7, ragged=True),
    "o2_neg0": dict(d=128, N=40, nw=5, L=20, W=2, neg=0, seed=108, lr=0.05, lam=1.0, scale=0.5),
    "o2_d256": dict(d=256, N=40, nw=6, L=30, W=5, neg=5, seed=109, lr=0.02, lam=1.0, scale=0.3),
    "o2_truncate_10000": dict(d=4, N=30, nw=1, L=10050, W=1, neg=1, seed=110, lr=0.05, lam=1.0, scale=1.0),
    "o2_empty_and_single": dict(d=128, N=20, nw=6, L=3, W=2, neg=2, seed=111, lr=0.05, lam=1.0, scale=0.5,
                                lens=[0, 1, 2, 3, 0, 3]),
}

O1_CASES = {
    "o1_d128": dict(d=128, N=100, E=300, neg=5, seed=200, lr=0.1, scale=0.3),
    "o1_d128_init_like_model": dict(d=128, N=100, E=200, neg=4, seed=201, lr=0.1, scale=1.0),
    "o1_d2": dict(d=2, N=34, E=156, neg=4, seed=202, lr=0.1, scale=1.0),
    "o1_d100_selfloops": dict(d=100, N=30, E=120, neg=3, seed=203, lr=0.05, scale=0.5, selfloop_every=9),
    "o1_neg0": dict(d=64, N=30, E=50, neg=0, seed=204, lr=0.05, scale=0.5),
}

O3_CASES = {
    "o3_d128_k5_dense": dict(d=128, N=120, K=5, seed=300, beta=0.1, lr=0.1, iters=1, onehot=False),
    "o3_d128_k4_onehot_iter5": dict(d=128, N=150, K=4, seed=301, beta=0.01, lr=0.1, iters=5, onehot=True),
    "o3_d2_k2": dict(d=2, N=34, K=2, seed=302, beta=0.01, lr=0.1, iters=5, onehot=False),
    "o3_d16_k3_subset_clip": dict(d=16, N=90, K=3, seed=303, beta=50.0, lr=0.1, iters=2, onehot=False, subset=True,
                                  cov_scale=0.01),
}


# legacy fused pass (stale train_sg): per pair o3 gradient + SGNS + combined write
SG_CASES = {
    "sg_d128_k3_ctx": dict(d=128, N=40, K=3, nw=4, L=12, W=2, neg=5, seed=500, lr=0.025, l1=1.0, l2=0.1, isnode=0),
    "sg_d128_k3_nodeemb": dict(d=128, N=40, K=3, nw=3, L=12, W=2, neg=5, seed=501, lr=0.025, l1=1.0, l2=0.1, isnode=1),
    "sg_d16_k4_w5": dict(d=16, N=40, K=4, nw=3, L=30, W=5, neg=3, seed=502, lr=0.05, l1=0.7, l2=0.2, isnode=0),
    "sg_d2_k2_karate": dict(d=2, N=34, K=2, nw=5, L=20, W=3, neg=4, seed=503, lr=0.1, l1=1.0, l2=0.1, isnode=0),
    "sg_d100_l2zero_w1": dict(d=100, N=40, K=3, nw=3, L=12, W=1, neg=5, seed=504, lr=0.025, l1=1.0, l2=0.0, isnode=0),
    "sg_d64_onehot_none": dict(d=64, N=50, K=5, nw=4, L=16, W=3, neg=4, seed=505, lr=0.05, l1=1.0, l2=0.5, isnode=0,
                               onehot=True, none_every=5),
}


# Python-twin fallback (utils/embedding.py:15-98): same input recipe as SG_CASES, different semantics
TWIN_CASES = {
    "twin_d128_ctx": dict(d=128, N=40, K=3, nw=3, L=10, W=2, neg=5, seed=600, lr=0.025, l1=1.0, l2=0.1, isnode=0),
    "twin_d16_k4": dict(d=16, N=30, K=4, nw=3, L=20, W=4, neg=3, seed=601, lr=0.05, l1=0.7, l2=0.2, isnode=0),
    "twin_d2_l2zero": dict(d=2, N=34, K=2, nw=4, L=20, W=3, neg=4, seed=602, lr=0.1, l1=1.0, l2=0.0, isnode=0),
    "twin_d64_nodeemb_none": dict(d=64, N=30, K=3, nw=3, L=12, W=3, neg=4, seed=603, lr=0.05, l1=1.0, l2=0.3, isnode=1,
                                  none_every=5),
    "twin_tiny_table_duplicates": dict(d=8, N=6, K=2, nw=3, L=12, W=2, neg=5, seed=604, lr=0.1, l1=1.0, l2=0.0,
                                       isnode=0),
}


def sg_inputs(c):
    rs = np.random.RandomState(c["seed"])
    d, N, K = c["d"], c["N"], c["K"]
    node = (rs.uniform(-1, 1, (N, d)) * 0.5).astype(np.float32)
    ctx = (rs.uniform(-1, 1, (N, d)) * 0.5).astype(np.float32)
    table = make_table(rs, N)
    mu = rs.uniform(-0.5, 0.5, (K, d)).astype(np.float32)
    inv = (rs.normal(size=(K, d, d)) * 0.3).astype(np.float32)  # deliberately not symmetric (column-major read!)
    if c.get("onehot"):
        pi = np.zeros((N, K), np.float32)
        pi[np.arange(N), rs.randint(0, K, size=N)] = 1.0
    else:
        p = rs.uniform(0, 1, (N, K)) ** 3
        pi = (p / p.sum(1, keepdims=True)).astype(np.float32)
    walks = []
    for w in range(c["nw"]):
        t = rs.randint(0, N, size=c["L"]).astype(np.uint32)
        if c.get("none_every"):
            t[::c["none_every"]] = TOKEN_NONE
        walks.append(t)
    return node, ctx, table, mu, inv, pi, walks


def sg_draws(rs, walks, window):
    """np.random consumption of the legacy train_sg per call: the LCG seed (2 draws), then one randint(window) per
    non-None token when window > 1 (old-pyx:343, 356-364).  Returns (seeds uint64[nw], reduced windows flat int32)."""
    seeds, rws = [], []
    for w in walks:
        r = rs.randint(0, 2 ** 24, size=2).astype(np.uint64)
        seeds.append((r[0] << np.uint64(24)) + r[1])
        rw = np.zeros(len(w), np.int32)
        if window > 1:
            for i, t in enumerate(w):
                if int(t) != TOKEN_NONE:
                    rw[i] = rs.randint(window)
        rws.append(rw)
    return np.asarray(seeds, np.uint64), (np.concatenate(rws) if rws else np.zeros(0, np.int32))


def make_table(rs, N, size=1000):
    """A small negative table with the reference's quirk: values in 1..N-1 (node ids used as rows), monotone."""
    return np.sort(rs.randint(1, N, size=size)).astype(np.uint32)


def o2_inputs(c):
    rs = np.random.RandomState(c["seed"])
    d, N = c["d"], c["N"]
    node = (rs.uniform(-1, 1, (N, d)) * c["scale"]).astype(np.float32)
    if c.get("ctx_zero"):
        ctx = np.zeros((N, d), np.float32)
    else:
        ctx = (rs.uniform(-1, 1, (N, d)) * c["scale"]).astype(np.float32)
    table = make_table(rs, N)
    walks = []
    for w in range(c["nw"]):
        if "lens" in c:
            L = c["lens"][w]
        elif c.get("ragged"):
            L = int(rs.randint(1, c["L"] + 1))
        else:
            L = c["L"]
        t = rs.randint(0, N, size=L).astype(np.uint32)
        if c.get("none_every"):
            t[::c["none_every"]] = TOKEN_NONE
        walks.append(t)
    return node, ctx, table, walks


def flatten_walks(walks):
    off = np.zeros(len(walks) + 1, np.int64)
    off[1:] = np.cumsum([len(w) for w in walks])
    flat = np.concatenate(walks).astype(np.uint32) if off[-1] else np.zeros(0, np.uint32)
    return np.ascontiguousarray(flat), off


def o1_inputs(c):
    rs = np.random.RandomState(c["seed"])
    d, N = c["d"], c["N"]
    node = (rs.uniform(-1, 1, (N, d)) * c["scale"]).astype(np.float32)
    table = make_table(rs, N)
    edges = rs.randint(0, N, size=(c["E"], 2)).astype(np.uint32)
    if c.get("selfloop_every"):
        edges[::c["selfloop_every"], 1] = edges[::c["selfloop_every"], 0]
    return node, table, edges


def o3_inputs(c):
    rs = np.random.RandomState(c["seed"])
    d, N, K = c["d"], c["N"], c["K"]
    node = rs.uniform(-1, 1, (N, d)).astype(np.float32)
    mu = rs.uniform(-0.5, 0.5, (K, d)).astype(np.float32)
    inv = np.empty((K, d, d), np.float32)
    for k in range(K):
        a = rs.normal(size=(d, d + 8))
        cov = a @ a.T / (d + 8) * c.get("cov_scale", 1.0) + 1e-3 * np.eye(d)
        m = np.linalg.inv(cov)
        m = m + 0.05 * rs.normal(size=(d, d)) * np.abs(m).mean()  # deliberately NOT symmetric
        inv[k] = m.astype(np.float32)
    if c["onehot"]:
        pi = np.zeros((N, K), np.float32)
        pi[np.arange(N), rs.randint(0, K, size=N)] = 1.0
    else:
        p = rs.uniform(0, 1, (N, K)) ** 4
        pi = (p / p.sum(1, keepdims=True)).astype(np.float32)
    if c.get("subset"):
        rows = np.sort(rs.choice(N, size=N // 2, replace=False)).astype(np.uint32)
    else:
        rows = np.arange(N, dtype=np.uint32)
    return node, mu, inv, pi, rows
