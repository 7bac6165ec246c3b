"""N>1 host logic on CPU: two ranks over torch.distributed (gloo, 127.0.0.1) build their block-cyclic shares in
structure-only mode, exchange mirror tables through the same PanguLU.connect() the GPU path uses, and check that
every remote block a rank reads is pushed by exactly one peer (its owner)."""
import os
import subprocess
import sys
import textwrap

import pytest

from helpers import ROOT

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from pangulu_b200 import PanguLU, matrices
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    s = PanguLU(matrices.poisson3d(10), nb=32, host_only=True, rank=rank, nproc=world)
    table = s.mirror_table()
    made = s.connect()
    p, q = int(s.info("p")), int(s.info("q"))
    # every mirrored block belongs to another rank of my process row (L) / process column (U) / both (diagonal halves)
    for brow, bcol, kind, voff, foff, doff, grp in table:
        owner = (brow %% p) * q + (bcol %% q)
        assert owner != rank
        if kind == 0 and brow > bcol: assert brow %% p == rank // q
        if kind == 0 and brow < bcol: assert bcol %% q == rank %% q
    gathered = [None] * world
    dist.all_gather_object(gathered, (len(table), made, [tuple(int(x) for x in r[:3]) for r in table]))
    if rank == 0:
        want_pushes = sum(g[0] for g in gathered)      # every mirrored block must be pushed once, by its owner
        got_pushes = sum(g[1] for g in gathered)
        assert want_pushes == got_pushes, (want_pushes, got_pushes)
        print(json.dumps({"mirrored": want_pushes, "pushes": got_pushes}))
    # tasks reference only blocks the rank holds (build_schedule would have aborted otherwise)
    tasks, wave_ptr = s.schedule()
    assert len(tasks) > 0
    s.finalize()
    dist.barrier()
    dist.destroy_process_group()
""") % ROOT


@pytest.mark.parametrize("world", [2, 4, 8])
def test_mirror_tables_and_pushes_over_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 1000 + world)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), PORT=port, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert '"pushes"' in outs[0]
