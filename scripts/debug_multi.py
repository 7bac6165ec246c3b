import os, sys, json, time
from types import SimpleNamespace
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pangulu_b200 import PanguLU, matrices
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29611", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
gl = dist.new_group(backend="gloo")
A = matrices.poisson3d(12); nb = 32
s = PanguLU(A, nb=nb, mode=1, rank=rank, nproc=world)
s.connect()
from oracle.oracle import Oracle
ref = PanguLU(A, nb=nb, host_only=True)
want = [SimpleNamespace(**vars(b)) for b in ref.blocks(copy=True)]
Oracle("r64").block_lu(nb, want)
index = {(b.brow, b.bcol, b.is_upper): b for b in want}
def check(tag):
    bad = 0; worst = 0
    for g in s.blocks(copy=True):
        w = index[(g.brow, g.bcol, g.is_upper)]
        if w.value.size:
            e = float(np.abs(g.value - w.value).max() / max(np.abs(w.value).max(), 1e-300))
            worst = max(worst, e if np.isfinite(e) else 1e30); bad += e > 1e-12 or not np.isfinite(e)
    print(f"rank {rank} {tag}: bad blocks {bad} worst {worst:.3e}", flush=True)
mode = os.environ.get("DBG", "a")
for run in range(3):
    if run:
        s.reset_values(True)
        torch.cuda.synchronize()
        if mode == "b": dist.barrier(group=gl)
    s.gstrf(); s.download(); check(f"run{run}")
    if mode == "b": dist.barrier(group=gl)
dist.barrier(); dist.destroy_process_group()
