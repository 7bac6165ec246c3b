"""Host logic of the driver (no GPU): reordering, symbolic pattern, block storage, schedule and flop model,
checked bit for bit against the golden vectors from the unmodified reference and against live reference runs."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import GOLDEN
from oracle.ref_run import unflatten
from pangulu_b200 import PanguLU, _lib, matrices

GOLDEN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "pattern" not in p)


def _same_structure(mine, ref_blocks):
    assert len(mine) == len(ref_blocks)
    for a, b in zip(mine, ref_blocks):
        assert (a.brow, a.bcol, a.is_upper) == (b.brow, b.bcol, b.is_upper)
        assert np.array_equal(a.colptr, b.colptr) and np.array_equal(a.rowidx, b.rowidx)
        assert np.array_equal(a.value, b.value)  # A's values scattered into the zero-filled pattern


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_structure_perm_and_flops_match_golden(case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    n, nb = int(z["n"]), int(z["nb"])
    A = sp.csc_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
    s = PanguLU(A, nb=nb, host_only=True)
    try:
        assert int(s.info("symbolic_nnz")) == int(z["symbolic_nnz"])
        assert np.array_equal(s.perms()["metis_perm"], z["metis_perm"])  # permutation indices: bit-exact
        _same_structure(s.blocks(), unflatten(z, "pre_", nb))
        assert int(s.info("flop")) == int(z["flop"])  # closed-form flop model == reference's merge-based counter
    finally:
        s.finalize()


@pytest.mark.parametrize("make,nb", [(lambda: matrices.poisson3d(14), 64), (lambda: matrices.laplace2d(60), 48),
                                     (lambda: matrices.stencil27(9), 128), (lambda: matrices.circuit_like(800, dmax=30, dtype=np.float64), 40)],
                         ids=["poisson3d-14", "laplace2d-60", "stencil27-9", "circuit-800"])
def test_structure_matches_live_reference(make, nb):
    from oracle.ref import ref_available
    from oracle.ref_run import run_reference

    if not ref_available("r64"):
        pytest.skip("oracle/_ref not built on this box")
    A = make()
    z = run_reference(A, nb, perf=True, solve=False)
    s = PanguLU(A, nb=nb, host_only=True)
    try:
        assert np.array_equal(s.perms()["metis_perm"], z["metis_perm"])
        _same_structure(s.blocks(), unflatten(z, "pre_", nb))
        assert int(s.info("flop")) == int(z["flop"])
        # CSR mirrors of strictly-lower blocks equal the reference routine's output
        from oracle.ref_run import add_csr_mirror

        for b in s.blocks():
            if b.brow > b.bcol:
                from types import SimpleNamespace

                m = add_csr_mirror(SimpleNamespace(colptr=b.colptr, rowidx=b.rowidx), nb)
                assert np.array_equal(m.rowptr, b.rowptr) and np.array_equal(m.colidx, b.colidx) and np.array_equal(m.csr2csc, b.csr2csc)
    finally:
        s.finalize()


@pytest.mark.parametrize("policy", [0, 1, 2], ids=["lazy", "eager", "windowed"])
def test_schedule_respects_dependencies(policy):
    """Every task of the static schedule runs in a later wave than the tasks that produce its operands, and no
    update of a destination runs after the wave that finalises it (lazy policy: exactly in that wave)."""
    A = matrices.poisson3d(12)
    nb = 32
    s = PanguLU(A, nb=nb, host_only=True, schedule=policy, split_diag=0, lookahead=6)
    try:
        tasks, wave_ptr = s.schedule()
        wave_of_task = np.repeat(np.arange(len(wave_ptr) - 1), np.diff(wave_ptr.astype(np.int64)))
        final = {}  # slot address -> wave in which the block is finalised (diag: keyed by the lower half)
        slot_addr = {}
        for i in range(int(s.info("num_slots"))):
            p = s.slot_ptr(i)
            sl = p.contents
            slot_addr[(sl.brow_pos, sl.bcol_pos, sl.is_upper)] = C_addr(p)
        for t, w in zip(tasks, wave_of_task):
            if t["kernel_id"] != _lib.TASK_SSSSM:
                final[int(t["opdst"])] = w
        n_upd = 0
        for t, w in zip(tasks, wave_of_task):
            k = int(t["task_level"])
            diag_lower = slot_addr[(k, k, 0)]
            if t["kernel_id"] in (_lib.TASK_TSTRF, _lib.TASK_GESSM):
                assert final[diag_lower] < w
            if t["kernel_id"] == _lib.TASK_SSSSM:
                assert final[int(t["op1"])] < w and final[int(t["op2"])] < w
                ready = 1 + max(final[int(t["op1"])], final[int(t["op2"])])
                if policy == 0:
                    assert final[int(t["opdst"])] == w  # applied in the wave that finalises the destination
                elif policy == 1:
                    assert w <= final[int(t["opdst"])] and w == ready  # as early as possible
                else:
                    assert w == max(ready, final[int(t["opdst"])] - 6)  # windowed, default lookahead
                n_upd += 1
        assert n_upd == int(s.info("n_ssssm")) and n_upd > 0
        # within a wave: SSSSM first, then GETRF, then TSTRF/GESSM
        for w in range(len(wave_ptr) - 1):
            ids = tasks["kernel_id"][int(wave_ptr[w]):int(wave_ptr[w + 1])]
            cls = np.where(ids == _lib.TASK_SSSSM, 0, np.where(ids == _lib.TASK_GETRF, 1, 2))
            assert np.all(np.diff(cls) >= 0)
    finally:
        s.finalize()


def C_addr(ptr):
    import ctypes

    return ctypes.addressof(ptr.contents)


@pytest.mark.parametrize("split", [0, 1], ids=["whole-diagonals", "diagonal-groups"])
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_block_cyclic_ownership_partitions_the_tasks(nproc, split):
    """2D block-cyclic owner-computes (src/pangulu_common.h:135, src/pangulu.c:83-90): the ranks' task lists are a
    partition of the single-rank list (GETRF tasks per independent group when the diagonal blocks are split: every rank
    derives the same groups from the symbolic pattern), and every operand a rank reads is a block it holds."""
    A = matrices.poisson3d(10)
    nb = 32
    key = lambda t: (int(t["row"]), int(t["col"]), int(t["kernel_id"]), int(t["task_level"]), int(t["compare_flag"]) >> 32)
    waves = lambda wp: np.repeat(np.arange(len(wp) - 1), np.diff(wp.astype(np.int64)))
    base = PanguLU(A, nb=nb, host_only=True, split_diag=split)
    t0, wp0 = base.schedule()
    all_tasks = sorted(key(t) + (int(w),) for t, w in zip(t0, waves(wp0)))
    base.finalize()
    got = []
    pq = {2: (1, 2), 4: (2, 2), 8: (2, 4)}[nproc]
    for r in range(nproc):
        s = PanguLU(A, nb=nb, host_only=True, rank=r, nproc=nproc, split_diag=split)
        assert (int(s.info("p")), int(s.info("q"))) == pq
        t, wp = s.schedule()
        for x in t:
            assert (int(x["row"]) % pq[0]) * pq[1] + int(x["col"]) % pq[1] == r
        # the wave of a task is part of the key: every rank places a task in the wave the single-rank schedule gives it, so the
        # waves are ONE global order in which every dependency -- local or on another rank -- points backwards.  That is what
        # keeps the persistent executors of several ranks, each fetching its own list in order, free of circular waits.
        got += [key(x) + (int(w),) for x, w in zip(t, waves(wp))]
        s.finalize()
    assert sorted(got) == all_tasks


def test_diagonal_groups_schedule():
    """Independent groups of diagonal blocks (schedule.cpp: diag_groups): the groups partition the block, the filled
    pattern is block diagonal over them, every group has its own GETRF task, the total work is unchanged, every task
    still runs after what it depends on, and the dependency depth shrinks."""
    A = matrices.poisson3d(14)
    nb = 32
    whole = PanguLU(A, nb=nb, host_only=True, split_diag=0)
    split = PanguLU(A, nb=nb, host_only=True, split_diag=1)
    try:
        assert split.info("flop") == whole.info("flop")
        assert split.info("n_ssssm") == whole.info("n_ssssm") and split.info("n_tstrf") == whole.info("n_tstrf")
        assert split.info("n_split_diag") > 0 and split.info("nwave") < whole.info("nwave")
        bl = int(split.info("block_length"))
        groups = {k: split.diag_groups(k) for k in range(bl)}
        lower = {b.brow: b for b in split.blocks(copy=True) if b.brow == b.bcol and not b.is_upper}
        offdiag = {(b.brow, b.bcol): b for b in split.blocks(copy=True) if b.brow != b.bcol}
        ngetrf = 0
        for k, bd in groups.items():
            assert bd[0] == 0 and bd[-1] == nb and all(x < y for x, y in zip(bd, bd[1:])) and len(bd) <= 5
            ngetrf += len(bd) - 1
            grp = np.searchsorted(np.array(bd), np.arange(nb), side="right") - 1
            b = lower[k]
            cols = np.repeat(np.arange(nb), np.diff(b.colptr.astype(np.int64)))
            assert np.array_equal(grp[b.rowidx], grp[cols]), f"diagonal block {k} is not block diagonal over its groups"
        assert ngetrf == int(split.info("n_getrf"))
        tasks, wave_ptr = split.schedule()
        wave = np.repeat(np.arange(len(wave_ptr) - 1), np.diff(wave_ptr.astype(np.int64)))
        getrf_wave = {}
        for t, w in zip(tasks, wave):
            if t["kernel_id"] == _lib.TASK_GETRF:
                getrf_wave[(int(t["row"]), int(t["compare_flag"]) >> 32)] = w
        assert len(getrf_wave) == ngetrf
        grp_of = lambda k, idx: np.unique(np.searchsorted(np.array(groups[k]), idx, side="right") - 1)
        checked = 0
        for t, w in zip(tasks, wave):
            kid, i, j, k = int(t["kernel_id"]), int(t["row"]), int(t["col"]), int(t["task_level"])
            if kid == _lib.TASK_TSTRF:       # block (i,k): groups of k in which it has columns
                b = offdiag[(i, j)]
                touched = grp_of(j, np.nonzero(np.diff(b.colptr.astype(np.int64)))[0])
                assert all(getrf_wave[(j, int(g))] < w for g in touched)
                checked += 1
            elif kid == _lib.TASK_GESSM:     # block (k,j): groups of k in which it has rows
                b = offdiag[(i, j)]
                touched = grp_of(i, np.unique(b.rowidx))
                assert all(getrf_wave[(i, int(g))] < w for g in touched)
                checked += 1
            elif kid == _lib.TASK_SSSSM and i == j:   # update of diagonal block i from level k: groups in which L(i,k) has rows
                touched = grp_of(i, np.unique(offdiag[(i, k)].rowidx))
                assert all(w <= getrf_wave[(i, int(g))] for g in touched)
                checked += 1
        assert checked > 0
    finally:
        whole.finalize()
        split.finalize()


def test_factor_file_round_trip(tmp_path):
    """Factor files (PanguLU's packed block format behind a small header): values written from one handle come back
    bit for bit in another handle of the same matrix; the file is parsed here independently from the layout description."""
    import struct

    from pangulu_b200 import PanguLU, matrices

    A = matrices.poisson3d(7)
    nb = 24
    s = PanguLU(A, nb=nb, host_only=True)
    rng = np.random.default_rng(9)
    vals = []
    for b in s.blocks(copy=False):
        b.value[:] = rng.uniform(-1, 1, b.value.shape)
        vals.append(b.value.copy())
    path = tmp_path / "factors.pglu"
    nbytes = s.save_factors(path)
    assert nbytes == os.path.getsize(path)
    # independent parse: header, then one packed block after the other
    raw = open(path, "rb").read()
    magic, version, vtype, szv, fnb, n, bl, nblocks = struct.unpack_from("<8sIIIIIIQ", raw, 0)
    assert magic == b"PGLUF01\x00" and version == 1 and vtype == 0 and szv == 8 and fnb == nb and n == A.shape[0] and nblocks == len(vals)
    off = struct.calcsize("<8sIIIIIIQ") + n * (3 * 4 + 2 * 8)
    for b, v in zip(s.blocks(copy=True), vals):
        nnz, brow, bcol, is_upper = struct.unpack_from("<QIII", raw, off)
        assert (nnz, brow, bcol) == (len(v), b.brow, b.bcol)
        off += 32
        assert np.array_equal(np.frombuffer(raw, np.float64, nnz, off), v)
        off += 8 * nnz
        assert np.array_equal(np.frombuffer(raw, np.uint32, nb + 1, off), b.colptr)
        off += 4 * (nb + 1)
        assert np.array_equal(np.frombuffer(raw, np.uint16, nnz, off), b.rowidx)
        off += 2 * nnz
        off += (8 - (4 * (nb + 1) + 2 * nnz) % 8) % 8
        if brow > bcol:
            assert np.array_equal(np.frombuffer(raw, np.uint32, nnz, off), b.csr2csc)
            off += 4 * nnz + 4 * (nb + 1) + 2 * nnz
            off += (8 - (4 * nnz + 4 * (nb + 1) + 2 * nnz) % 8) % 8
    assert off == len(raw)
    t = PanguLU(A, nb=nb, host_only=True)
    assert t.load_factors(path) == len(vals)
    for b, v in zip(t.blocks(copy=True), vals):
        assert np.array_equal(b.value, v)
    s.finalize(); t.finalize()
