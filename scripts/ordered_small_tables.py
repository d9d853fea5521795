"""ORDERED o2 variants on small tables (many equal samples / hazards): which kernel should the launcher pick?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comemb_b200.utils.training_sdg_inner as K
from comemb_b200 import _lib

K.init()
lib = _lib.load()
d, L, nw = 128, 80, 100
for N in (34, 200, 1000, 5000, 20000):
    g = torch.Generator(device='cuda').manual_seed(0)
    node0 = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
    ctx0 = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
    table = torch.randint(0, N, (1000000,), device='cuda', generator=g, dtype=torch.int32)
    walks = torch.randint(0, N, (nw * L,), device='cuda', generator=g, dtype=torch.int32)
    off = torch.arange(nw + 1, device='cuda', dtype=torch.int64) * L
    out = []
    for variant in (800, 700, 0):
        _lib.check(lib.comemb_set_tuning(0, 0, variant))
        for rep in range(2):
            node, ctx = node0.clone(), ctx0.clone()
            torch.cuda.synchronize(); t = time.perf_counter()
            K.o2_batch(node, ctx, walks, off, None, 0.025, 5, 10, table, mode=K.MODE_ORDERED, base_seed=1)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        out.append(nw * 1490 / dt)
    _lib.check(lib.comemb_set_tuning(0, 0, 0))
    print("N=%6d  plain %.3g  pipelined %.3g  team %.3g pair-updates/s" % (N, *out), flush=True)
