"""One small ORDERED o2 launch (d=128, 5 negatives) for `ncu --set full --import-source on -k regex:ordered`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comemb_b200.utils.training_sdg_inner as K

K.init()
N, d, L, nw = 100000, 128, 80, 40
g = torch.Generator(device='cuda').manual_seed(0)
node = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
ctx = (torch.rand((N, d), device='cuda', generator=g) - 0.5) * 0.1
table = torch.randint(1, N, (5000000,), device='cuda', generator=g, dtype=torch.int32)
walks = torch.randint(0, N, (nw * L,), device='cuda', generator=g, dtype=torch.int32)
off = torch.arange(nw + 1, device='cuda', dtype=torch.int64) * L
K.o2_batch(node, ctx, walks, off, None, 0.025, 5, 10, table, mode=K.MODE_ORDERED, base_seed=1)
torch.cuda.synchronize()
print("ok")
