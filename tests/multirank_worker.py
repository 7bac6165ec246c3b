"""Worker of the multi-rank GPU tests (one process per rank; launched by tests/test_multigpu_gpu.py and
tests/test_production_gpu.py).  Factors one matrix on WORLD_SIZE ranks -- blocks 2D block-cyclic over the ranks,
finished blocks pushed GPU to GPU inside the persistent kernel -- and compares the blocks THIS rank owns with the
expected single-rank factors (EXPECT = npz written by oracle.ref_run, i.e. the unmodified reference; else the C
oracle's block LU), then runs the distributed self-check ||L(U1) - A1|| / ||A1||.

One rank per GPU when the box has WORLD_SIZE GPUs (NCCL bootstrap).  With fewer GPUs the ranks share them
(rank % device_count, gloo bootstrap): the persistent kernels of the ranks are then time-sliced by the driver, which
is slow but exercises exactly the same exchange (CUDA-IPC mappings, in-kernel pushes, system-scope flags)."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pangulu_b200 import PanguLU, matrices  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
dev = rank % ndev
torch.cuda.set_device(dev)
shared = ndev < world
if shared:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=world)
else:
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=world,
                            device_id=torch.device("cuda", dev))
gen = {"poisson3d": matrices.poisson3d, "laplace2d": matrices.laplace2d, "stencil27": matrices.stencil27}[os.environ.get("KIND", "poisson3d")]
A = gen(int(os.environ["GRID"]))
nb = int(os.environ["NB"])
tol = float(os.environ.get("TOL", "1e-12"))
s = PanguLU(A, nb=nb, mode=1, rank=rank, nproc=world)
s.connect()
s.gstrf()
s.download()
got = s.blocks(copy=True)          # owned blocks only

expect = os.environ.get("EXPECT", "")
if expect:
    from oracle.ref_run import unflatten

    with np.load(expect, allow_pickle=False) as z:
        want = unflatten(z, "post_", nb)
        assert np.array_equal(s.perms()["metis_perm"], z["metis_perm"]), "permutation differs from the reference's"
    source = "reference"
else:
    from oracle.oracle import Oracle

    ref = PanguLU(A, nb=nb, host_only=True)
    want = [SimpleNamespace(**vars(b)) for b in ref.blocks(copy=True)]
    ref.finalize()
    Oracle("r64").block_lu(nb, want)
    source = "oracle"
index = {(b.brow, b.bcol, b.is_upper): b for b in want}


def compare(blocks):
    worst, bad = 0.0, []
    for g in blocks:
        w = index[(g.brow, g.bcol, g.is_upper)]
        assert np.array_equal(g.rowidx, w.rowidx) and np.array_equal(g.colptr, w.colptr)     # structure: bit-exact
        if w.value.size:
            assert np.isfinite(g.value).all(), ("non-finite values in block", g.brow, g.bcol, g.is_upper)
            e = float(np.abs(g.value - w.value).max() / max(np.abs(w.value).max(), 1e-300))
            if e > tol:
                bad.append((g.brow, g.bcol, g.is_upper, e))
            worst = max(worst, e)
    return worst, bad


worst, bad = compare(got)
assert not bad, (rank, bad[:8], len(bad))
check = s.factor_check_distributed()
assert check < 1e-12, ("distributed L(U1) check", check)
s.reset_values(True); s.gstrf(); s.download()   # a second run on the same wiring must work too
worst2, bad2 = compare(s.blocks(copy=True))
assert not bad2, (rank, "second run", bad2[:8], len(bad2))
print(json.dumps({"rank": rank, "world": world, "gpus": ndev, "shared_gpus": shared, "owned_blocks": len(got), "worst": max(worst, worst2),
                  "factor_check": check, "expected_from": source, "gstrf_s": s.info("time_gstrf")}))
s.finalize()
dist.barrier()
dist.destroy_process_group()
