"""Multi-rank gstrf (2D block-cyclic ownership, blocks pushed GPU to GPU inside the persistent kernel): every rank's
owned blocks must equal the single-rank oracle's, structure bit-exact and values within 1e-12, and the distributed
self-check must hold.  One rank per GPU when the box has enough GPUs; otherwise the ranks share the GPUs that exist
(time-sliced) -- the tests run, they do not skip, on a one-GPU box."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

WORKER = os.path.join(ROOT, "tests", "multirank_worker.py")


def run_ranks(world, env_extra, timeout=600):
    port = str(29700 + os.getpid() % 200 + world)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), PORT=port, OMP_NUM_THREADS="4", **env_extra)
        procs.append(subprocess.Popen([sys.executable, WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=timeout)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return outs


@pytest.mark.parametrize("world,grid,nb", [(2, 12, 32), (2, 16, 64), (4, 14, 32), (8, 16, 32)])
def test_multi_gpu_factors_match_single_rank_oracle(world, grid, nb):
    run_ranks(world, {"GRID": str(grid), "NB": str(nb)})
