import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comemb_b200.utils.training_sdg_inner as K
K.init()
N, d, L, nw = 100000, 128, 80, 400
g = torch.Generator(device='cuda').manual_seed(0)
node = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
ctx = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
table = torch.randint(1, N, (5000000,), device='cuda', generator=g, dtype=torch.int32)
walks = torch.randint(0, N, (nw * L,), device='cuda', generator=g, dtype=torch.int32)
off = torch.arange(nw + 1, device='cuda', dtype=torch.int64) * L
for mode, name in ((K.MODE_ORDERED, 'ordered'),):
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        K.o2_batch(node, ctx, walks, off, None, 0.025, 5, 10, table, mode=mode, base_seed=1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(name, '%.3g pairs/s' % (nw * 1490 / dt), dt)
