"""ORDERED-mode throughput (one warp replaying the reference's sequential stream) for o2 and o1 at d=128, for the
three kernel variants (8: single warp; 7: single warp, software-pipelined; 0: default = one warp per target row for
o2, pipelined for o1); also checks that all give the same bits."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comemb_b200.utils.training_sdg_inner as K
from comemb_b200 import _lib

K.init()
lib = _lib.load()
N, d, L, nw = 100000, 128, 80, 400
g = torch.Generator(device='cuda').manual_seed(0)
node0 = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
ctx0 = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
table = torch.randint(1, N, (5000000,), device='cuda', generator=g, dtype=torch.int32)
walks = torch.randint(0, N, (nw * L,), device='cuda', generator=g, dtype=torch.int32)
off = torch.arange(nw + 1, device='cuda', dtype=torch.int64) * L
edges = torch.randint(0, N, (200000, 2), device='cuda', generator=g, dtype=torch.int32)
res = {}
for variant, name in ((800, 'plain'), (700, 'pipelined'), (0, 'default')):
    _lib.check(lib.comemb_set_tuning(0, 0, variant))
    for rep in range(2):
        node, ctx = node0.clone(), ctx0.clone()
        torch.cuda.synchronize(); t = time.perf_counter()
        K.o2_batch(node, ctx, walks, off, None, 0.025, 5, 10, table, mode=K.MODE_ORDERED, base_seed=1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('o2 ordered %-9s %.3g pair-updates/s (%.3f s)' % (name, nw * 1490 / dt, dt), flush=True)
    res['o2', name] = (node, ctx)
    for rep in range(2):
        node = node0.clone()
        torch.cuda.synchronize(); t = time.perf_counter()
        K.o1_batch(node, edges, None, 0.025, 5, table, mode=K.MODE_ORDERED, base_seed=1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('o1 ordered %-9s %.3g directed updates/s (%.3f s)' % (name, 2 * edges.shape[0] / dt, dt), flush=True)
    res['o1', name] = (node,)
_lib.check(lib.comemb_set_tuning(0, 0, 0))
for k in ('o2', 'o1'):
    for v in ('pipelined', 'default'):
        same = all(torch.equal(a, b) for a, b in zip(res[k, 'plain'], res[k, v]))
        print(k, v, '== plain (bit-exact):', same)
        assert same
