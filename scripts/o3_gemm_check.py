"""Quick device check of the tcgen05 o3 path: default (grouped 3xTF32 GEMM) vs COMEMB_VARIANT_ROUND1 (fp64-pipe kernel) vs
the oracle, plus timings at the bench shape.  Run on a B200:  python scripts/o3_gemm_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comemb_b200.utils.training_sdg_inner as K  # noqa: E402
from comemb_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

K.init()
lib = _lib.load()


def run(variant, node, rows, mu, inv_t, comm, weight, beta, lr, iters):
    _lib.check(lib.comemb_set_tuning(0, 0, variant))
    try:
        dn = torch.from_numpy(node).cuda()
        K.o3_batch_top1(dn, rows, mu, inv_t, comm, weight, beta, lr, iters=iters)
        torch.cuda.synchronize()
        return dn.cpu().numpy()
    finally:
        _lib.check(lib.comemb_set_tuning(0, 0, 0))


rs = np.random.RandomState(3)
for N, Kc, iters in ((70, 1, 1), (3001, 7, 1), (3001, 7, 3), (20000, 50, 1)):
    d = 128
    node = (rs.uniform(-1, 1, (N, d)) * 0.3).astype(np.float32)
    mu = rs.uniform(-0.5, 0.5, (Kc, d)).astype(np.float32)
    inv = (rs.normal(size=(Kc, d, d)) * 0.05 + np.eye(d)).astype(np.float32)
    comm = rs.randint(0, Kc, size=N).astype(np.int32)
    weight = np.ones(N, np.float32)
    weight[::11] = rs.uniform(0.2, 0.99, size=weight[::11].size).astype(np.float32)
    comm[::17] = -1
    weight[::17] = 0
    dmu, dinv = torch.from_numpy(mu).cuda(), torch.from_numpy(inv).cuda()
    inv_t = K.transpose_blocks(dinv)
    dc, dw = torch.from_numpy(comm).cuda(), torch.from_numpy(weight).cuda()
    a = run(0, node, None, dmu, inv_t, dc, dw, 5.0, 0.05, iters)
    b = run(600, node, None, dmu, inv_t, dc, dw, 5.0, 0.05, iters)
    ua, ub = a - node, b - node
    print("N=%d K=%d iters=%d: max|upd gemm - upd fp64| = %.3e (max |upd| %.3e)  rel %.2e  moved rows %d/%d" % (
        N, Kc, iters, np.abs(ua - ub).max(), np.abs(ub).max(), np.abs(ua - ub).max() / np.abs(ub).max(),
        int((np.abs(ua).max(1) > 0).sum()), int((np.abs(ub).max(1) > 0).sum())), flush=True)

# timing at the bench shape
N, Kc, d = 100000, 50, 128
node = (rs.uniform(-1, 1, (N, d)) * 0.05).astype(np.float32)
dn = torch.from_numpy(node).cuda()
dmu = torch.from_numpy(rs.uniform(-0.5, 0.5, (Kc, d)).astype(np.float32)).cuda()
inv_t = K.transpose_blocks(torch.from_numpy((rs.normal(size=(Kc, d, d)) * 0.05 + np.eye(d)).astype(np.float32)).cuda())
dc = torch.from_numpy((np.arange(N) % Kc).astype(np.int32)).cuda()
dw = torch.ones(N, device="cuda")
for variant in (0, 600):
    _lib.check(lib.comemb_set_tuning(0, 0, variant))
    for _ in range(3):
        K.o3_batch_top1(dn, None, dmu, inv_t, dc, dw, 0.1, 0.025, iters=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        K.o3_batch_top1(dn, None, dmu, inv_t, dc, dw, 0.1, 0.025, iters=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("variant %d: %.3f ms per o3 step over %d rows = %.3e rows/s" % (variant, ms, N, N / ms * 1e3), flush=True)
_lib.check(lib.comemb_set_tuning(0, 0, 0))
