"""Does the benchmark configuration actually learn?  Train the bench's SBM workload (100K nodes, 50 blocks, d=128) with
the Hogwild o2 kernel at full GPU concurrency for a few passes and report community NMI (k-means on a node sample)
and node-classification micro-F1, for red.add and plain-store scatter."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import comemb_b200.utils.training_sdg_inner as K  # noqa: E402
from comemb_b200 import _lib  # noqa: E402
from comemb_b200.evaluation import community_nmi, node_classification_micro_f1  # noqa: E402
from comemb_b200.utils import graph_utils as gu  # noqa: E402

K.init()
G, block = bench.build_workload()
n, d, L, W, neg = 100000, 128, 80, 10, 5
deg = np.ascontiguousarray(np.diff(G.rowptr), np.float64)
table = torch.empty(5000000, dtype=torch.int32, device="cuda")
_lib.check(_lib.load().comemb_make_table(deg.ctypes.data, deg.size, 0.75, table.data_ptr(), table.numel(), None))
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sample = np.random.RandomState(0).choice(n, 20000, replace=False)
for tag, flags in (("red.add", K.F_ATOMIC), ("plain", 0)):
    node_h, ctx_h = bench.init_tables_host(n, d)
    node, ctx = torch.from_numpy(node_h).cuda(), torch.from_numpy(ctx_h).cuda()
    t0 = time.perf_counter()
    for p in range(passes):
        walks, lens = gu.build_deepwalk_corpus(G, passes, L, alpha=0.0, seed=5, mode=gu.MODE_HOGWILD,
                                               return_device=True, first_walk=p * n, n_out=n)
        off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
        K.o2_batch(node, ctx, walks.reshape(-1), off, None, 0.025, neg, W, table, mode=K.MODE_HOGWILD, flags=flags,
                   base_seed=11 + p)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    x = node.cpu().numpy()
    print("%-8s %d passes in %.2fs: NMI(kmeans, 20K sample) %.3f  micro-F1 %.3f  finite %s" % (
        tag, passes, dt, community_nmi(x[sample], block[sample], k=50, method="kmeans"),
        node_classification_micro_f1(x[sample], block[sample]), bool(np.isfinite(x).all())), flush=True)
