"""The two hand-written GMM kernels alone (N1): comemb_gmm_estep and comemb_gmm_mstep at N=100K, d=128, K=50.

Prints one JSON line: milliseconds per launch (CUDA events, best of 5 after warm-up), algorithmic TFLOP/s
(2*N*K*d^2 per kernel) and executed tensor TFLOP/s (3 TF32 products per fp32 product), against the TF32 dense peak
taken as half of MEASURED_PEAKS.json's sustained bf16 figure."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
from comemb_b200 import _lib  # noqa: E402


def best_ms(fn, reps=5, warm=10):
    for _ in range(warm):  # also lets the SM clock ramp up
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


if __name__ == "__main__":
    n, d, K = 100000, 128, 50
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n, d), device="cuda", generator=g)
    P = torch.triu(torch.randn((K, d, d), device="cuda", generator=g)) * 0.05 + torch.eye(d, device="cuda")
    mu = torch.randn((K, d), device="cuda", generator=g) * 0.5
    bias = torch.bmm(mu[:, None, :], P).reshape(K, d).contiguous()
    sq = torch.empty((n, K), device="cuda")
    dense = torch.rand((n, K), device="cuda", generator=g)
    dense /= dense.sum(1, keepdim=True)
    lab = torch.randint(0, K, (n,), device="cuda", generator=g)
    onehot = torch.zeros((n, K), device="cuda")
    onehot[torch.arange(n, device="cuda"), lab] = 1
    scat = torch.empty((K, d, d), device="cuda")
    lib = _lib.load()
    st = _lib.stream_ptr()
    flop = 2.0 * n * K * d * d
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] / 2
    except Exception:
        pass
    out = {"n": n, "d": d, "K": K, "tf32_peak_tflops": peak}
    for _ in range(300):  # ~1 s of work first: the SM clock has to ramp up before anything is timed
        _lib.check(lib.comemb_gmm_estep(x.data_ptr(), n, d, P.data_ptr(), bias.data_ptr(), K, sq.data_ptr(), st))
    torch.cuda.synchronize()
    ms = best_ms(lambda: _lib.check(lib.comemb_gmm_estep(x.data_ptr(), n, d, P.data_ptr(), bias.data_ptr(), K, sq.data_ptr(), st)))
    out["estep"] = {"ms": round(ms, 3), "algorithmic_tflops": round(flop / ms / 1e9, 1), "executed_tf32_tflops": round(3 * flop / ms / 1e9, 1)}
    for name, r in (("mstep_dense_resp", dense), ("mstep_onehot_resp", onehot)):
        ms = best_ms(lambda: _lib.check(lib.comemb_gmm_mstep(x.data_ptr(), n, d, r.data_ptr(), mu.data_ptr(), K, scat.data_ptr(), st)))
        out[name] = {"ms": round(ms, 3), "algorithmic_tflops": round(flop / ms / 1e9, 1),
                     "executed_tf32_tflops": round(3 * flop / ms / 1e9, 1)}
    diff = x[None] - mu[:, None]
    ms = best_ms(lambda: torch.bmm((diff * dense.T[:, :, None]).transpose(1, 2), diff), reps=3, warm=2)
    out["mstep_library_bmm_with_temporaries"] = {"ms": round(ms, 3)}
    print(json.dumps(out))
