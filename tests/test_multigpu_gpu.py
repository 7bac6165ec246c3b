"""Multi-GPU gstrf (one rank per GPU, 2D block-cyclic ownership, blocks pushed over NVLink inside the persistent
kernel): every rank's owned blocks must equal the single-rank oracle's.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys
import textwrap

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys, json
    from types import SimpleNamespace
    import numpy as np
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from pangulu_b200 import PanguLU, matrices
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    A = matrices.poisson3d(int(os.environ["GRID"]))
    nb = int(os.environ["NB"])
    s = PanguLU(A, nb=nb, mode=1, rank=rank, nproc=world)
    s.connect()
    s.gstrf()
    s.download()
    got = s.blocks(copy=True)          # owned blocks only
    # expected factors: single-process oracle (C restatement of the reference kernels)
    from oracle.oracle import Oracle
    ref = PanguLU(A, nb=nb, host_only=True)
    want = [SimpleNamespace(**vars(b)) for b in ref.blocks(copy=True)]
    Oracle("r64").block_lu(nb, want)
    index = {(b.brow, b.bcol, b.is_upper): b for b in want}
    def compare(blocks):
        worst, bad = 0.0, []
        for g in blocks:
            w = index[(g.brow, g.bcol, g.is_upper)]
            assert np.array_equal(g.rowidx, w.rowidx) and np.array_equal(g.colptr, w.colptr)
            if w.value.size:
                assert np.isfinite(g.value).all(), ("non-finite values in block", g.brow, g.bcol, g.is_upper)
                e = float(np.abs(g.value - w.value).max() / max(np.abs(w.value).max(), 1e-300))
                if e > 1e-12:
                    bad.append((g.brow, g.bcol, g.is_upper, e))
                worst = max(worst, e)
        return worst, bad
    worst, bad = compare(got)
    assert not bad, (rank, bad[:8], len(bad))
    s.reset_values(True); s.gstrf(); s.download()   # a second run on the same wiring must work too
    worst2, bad2 = compare(s.blocks(copy=True))
    assert not bad2, (rank, "second run", bad2[:8], len(bad2))
    print(json.dumps({"rank": rank, "owned_blocks": len(got), "worst": worst}))
    s.finalize(); ref.finalize()
    dist.barrier(); dist.destroy_process_group()
""") % ROOT


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,grid,nb", [(2, 12, 32), (2, 16, 64), (4, 14, 32), (8, 16, 32)])
def test_multi_gpu_factors_match_single_rank_oracle(world, grid, nb, tmp_path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29700 + os.getpid() % 200 + world)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), PORT=port, GRID=str(grid), NB=str(nb), OMP_NUM_THREADS="4")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=240)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
