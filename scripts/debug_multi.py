import os, sys, json, time
from types import SimpleNamespace
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pangulu_b200 import PanguLU, matrices
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29611", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
A = matrices.poisson3d(int(os.environ.get("GRID", "14"))); nb = int(os.environ.get("NB", "32"))
from oracle.oracle import Oracle
ref = PanguLU(A, nb=nb, host_only=True)
want = [SimpleNamespace(**vars(b)) for b in ref.blocks(copy=True)]
Oracle("r64").block_lu(nb, want)
index = {(b.brow, b.bcol, b.is_upper): b for b in want}
s = PanguLU(A, nb=nb, mode=1, rank=rank, nproc=world)
s.connect()
s.gstrf()
# fetch the WHOLE value region (owned + mirrors)
vb = int(s.info("value_bytes"))
sl0 = s.slot_ptr(0).contents
s.L.platform("memcpy")(sl0.value, sl0.d_value, vb, 1)
p, q = int(s.info("p")), int(s.info("q"))
bad_own, bad_mir = [], []
for g in s.blocks(copy=True, owned_only=False):
    w = index[(g.brow, g.bcol, g.is_upper)]
    if not w.value.size: continue
    e = float(np.abs(g.value - w.value).max() / max(np.abs(w.value).max(), 1e-300))
    if not (e <= 1e-12):
        (bad_own if g.owned else bad_mir).append((g.brow, g.bcol, g.is_upper, (g.brow % p) * q + g.bcol % q, float(f"{e:.2e}")))
print(f"rank {rank}: bad owned {len(bad_own)} {bad_own[:6]} | bad mirrored {len(bad_mir)} {bad_mir[:6]}", flush=True)
dist.barrier(); dist.destroy_process_group()
