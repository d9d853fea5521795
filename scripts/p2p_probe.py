"""Probe which peer-memory operations of the sharded o2 kernel work over NVLink (2 ranks)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import comemb_b200.utils.training_sdg_inner as K
from comemb_b200.sharded import ShardedTables
K.init()
n, d, L = 1000, 128, 20
st = ShardedTables(n, d)
rs = np.random.RandomState(0)
st.load_rows((rs.uniform(-1, 1, (n, d)) * 0.05).astype(np.float32), (rs.uniform(-1, 1, (n, d)) * 0.05).astype(np.float32))
dist.barrier()
table = torch.from_numpy(np.sort(rs.randint(1, n, 20000)).astype(np.int32)).cuda()
walks = torch.from_numpy(rs.randint(0, n, (1, L)).astype(np.int32)).cuda()
off = torch.tensor([0, L], dtype=torch.int64, device="cuda")
sd = torch.tensor([99], dtype=torch.int64, device="cuda")
if rank == 0:
    print("peer ptrs", [hex(p) for p in st._peer_ptrs[1]], flush=True)
    for lr in (0.0, 0.025):
        st.o2(walks[0], off, sd, lr, 5, 5, table)
        torch.cuda.synchronize(); print("sharded o2 lr=%g ok" % lr, flush=True)
dist.barrier()
dist.destroy_process_group()
