"""ORDERED o2 throughput: the dataflow replay (csrc/sgns_flow.cu) against the one-CTA replay, pair updates per second.

    python scripts/ordered_bench.py [--rows 100000] [--walks 20000] [--len 80]
Prints one JSON line per (variant, max_warps).  Both variants produce the same bits (tests/test_gpu_parity.py)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__  # noqa: E402,F401  (puts the package on the path)
from comemb_b200 import _lib  # noqa: E402
from comemb_b200.utils import training_sdg_inner as K  # noqa: E402


def run(rows, walks, length, window, neg, variant, max_warps, reps=2, team_walks=None):
    rng = np.random.RandomState(1)
    dev = torch.device("cuda:0")
    node = torch.from_numpy(((rng.rand(rows, 128) - 0.5) / 128).astype(np.float32)).to(dev)
    ctx = torch.zeros_like(node)
    table = torch.from_numpy(rng.randint(0, rows, size=1 << 22).astype(np.uint32)).to(dev)
    nw = team_walks if team_walks else walks
    flat = torch.from_numpy(rng.randint(0, rows, size=nw * length).astype(np.uint32)).to(dev)
    off = torch.arange(nw + 1, dtype=torch.int64, device=dev) * length
    pairs = nw * sum(min(length, i + window + 1) - max(0, i - window) - 1 for i in range(length))
    best = None
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with _lib.opts(variant=variant, max_warps=max_warps):
            e0.record()
            K.o2_batch(node, ctx, flat, off, None, 0.025, neg, window, table, mode=K.MODE_ORDERED, base_seed=r)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if r and (best is None or ms < best):
            best = ms
    return {"rows": rows, "walks": nw, "len": length, "window": window, "neg": neg, "variant": variant,
            "max_warps": max_warps, "pairs": pairs, "ms": round(best, 3), "pairs_per_s": round(pairs / best * 1e3)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, nargs="+", default=[1000, 10000, 100000, 1000000])
    ap.add_argument("--walks", type=int, default=20000)
    ap.add_argument("--len", type=int, default=80)
    ap.add_argument("--window", type=int, default=5)
    ap.add_argument("--neg", type=int, default=5)
    ap.add_argument("--warps", type=int, nargs="+", default=[0])
    a = ap.parse_args()
    for rows in a.rows:
        print(json.dumps(run(rows, a.walks, a.len, a.window, a.neg, _lib.VARIANT_ORDERED_TEAM, 0, reps=1,
                             team_walks=min(a.walks, 300))), flush=True)
        for mw in a.warps:
            print(json.dumps(run(rows, a.walks, a.len, a.window, a.neg, _lib.VARIANT_ORDERED_FLOW, mw)), flush=True)
