"""Per-kernel parity on the GPU: every block kernel is called through the platform C ABI
(pangulu_platform_0201001_{getrf,tstrf,gessm,ssssm,hybrid_batched}) on device slots and compared with the
CPU oracle (oracle/pangulu_oracle.c, itself pinned to the reference CPU platform by tests/test_oracle.py).

Tolerance: block-wise relative error <= 1e-12 (R64/CR64) or 1e-5 (R32/CR32) -- the north-star bound.
"""
import copy
import ctypes as C
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import TOL, DeviceBlocks, blocks_of, rel_err, trefethen_matrix
from pangulu_b200 import _lib, matrices

pytestmark = pytest.mark.gpu

CASES = [
    ("trefethen-nb4", lambda: trefethen_matrix(), 4),
    ("trefethen-nb10", lambda: trefethen_matrix(), 10),
    ("poisson3d-8-nb32", lambda: matrices.poisson3d(8), 32),
    ("laplace2d-24-nb64", lambda: matrices.laplace2d(24), 64),
    ("stencil27-6-nb48", lambda: matrices.stencil27(6), 48),
]


def _oracle(vtype):
    from oracle.oracle import Oracle

    return Oracle(vtype)


def _randomise(blocks, rng, dtype):
    """Replace the values by seeded random ones (diagonally dominant diagonal blocks) so that every structural
    entry -- including explicit zeros from fill -- carries data."""
    for b in blocks:
        v = rng.uniform(-1, 1, size=b.value.shape)
        if np.dtype(dtype).kind == "c":
            v = v + 1j * rng.uniform(-1, 1, size=b.value.shape)
        b.value = v.astype(dtype)
        if b.brow == b.bcol and b.is_upper:
            nb = len(b.colptr) - 1
            for r in range(nb):
                if b.colptr[r] < b.colptr[r + 1]:
                    b.value[b.colptr[r]] = 4.0 * nb  # the diagonal entry is first in each row


def _index(blocks):
    off = {(b.brow, b.bcol): b for b in blocks if b.brow != b.bcol}
    diag = {}
    for b in blocks:
        if b.brow == b.bcol:
            diag.setdefault(b.brow, [None, None])[1 if b.is_upper else 0] = b
    return off, diag


@pytest.fixture(params=[1, 2], ids=["ssssm-sparse", "ssssm-tile"])
def ssssm_mode(request):
    """Force the SSSSM path: 1 = sparse scatter kernel, 2 = dense-tile (DMMA / FMA tile) kernel."""
    yield request.param
    for vt in ("r64", "r32", "cr64", "cr32"):
        _lib.load(vt).pangulu_sm100_set_option(b"ssssm_mode", 0)


@pytest.mark.parametrize("vtype", ["r64", "r32", "cr64", "cr32"])
@pytest.mark.parametrize("name,make,nb", CASES, ids=[c[0] for c in CASES])
def test_single_kernels(name, make, nb, vtype, ssssm_mode):
    dtype = _lib.VALUE_DTYPES[vtype]
    _lib.load(vtype).pangulu_sm100_set_option(b"ssssm_mode", ssssm_mode)
    blocks = blocks_of(make(), nb, vtype)
    rng = np.random.default_rng(1234)
    _randomise(blocks, rng, dtype)
    off, diag = _index(blocks)
    O = _oracle(vtype)
    D = DeviceBlocks(vtype, nb)
    P = D.P
    tol = TOL[vtype]
    try:
        k = sorted(diag)[0]
        # --- GETRF on the first diagonal block
        lo, up = diag[k]
        want_lo, want_up = copy.deepcopy(lo), copy.deepcopy(up)
        O.getrf(nb, want_lo, want_up)
        s_lo, s_up = D.slot(lo), D.slot(up)
        D.link(s_lo, s_up)
        P("getrf")(nb, C.addressof(s_up), 0)
        P("synchronize")()
        got_lo, got_up = D.fetch(s_lo, len(lo.value)), D.fetch(s_up, len(up.value))
        assert rel_err(got_lo, want_lo.value) <= tol and rel_err(got_up, want_up.value) <= tol
        # eager download refreshed the host arrays too (reference semantics, pangulu_platform_0201000.cu:639-640)
        assert np.array_equal(lo.value, got_lo) and np.array_equal(up.value, got_up)
        # --- TSTRF / GESSM on every panel block of level k against the factored diagonal
        for (i, j), b in sorted(off.items()):
            if j == k and i > k:
                want = copy.deepcopy(b)
                O.tstrf(nb, want, want_up)
                sb = D.slot(b)
                P("tstrf")(nb, C.addressof(sb), C.addressof(s_lo), 0)  # either half may be passed
                P("synchronize")()
                assert rel_err(D.fetch(sb, len(b.value)), want.value) <= tol, f"tstrf {(i, j)}"
            if i == k and j > k:
                want = copy.deepcopy(b)
                O.gessm(nb, want, want_lo)
                sb = D.slot(b)
                P("gessm")(nb, C.addressof(sb), C.addressof(s_up), 0)
                P("synchronize")()
                assert rel_err(D.fetch(sb, len(b.value)), want.value) <= tol, f"gessm {(i, j)}"
        # --- SSSSM on every (i,j,k) triple of level k, off-diagonal and diagonal destinations
        rows = [i for (i, j) in off if j == k and i > k]
        cols = [j for (i, j) in off if i == k and j > k]
        n_done = 0
        for i in rows:
            for j in cols:
                a, b = off[(i, k)], off[(k, j)]
                sa, sb = D.slot(a), D.slot(b)
                if i == j:
                    dl, du = diag[i]
                    wl, wu = copy.deepcopy(dl), copy.deepcopy(du)
                    O.ssssm_diag(nb, wl, wu, a, b)
                    sl, su = D.slot(dl), D.slot(du)
                    D.link(sl, su)
                    P("ssssm")(nb, C.addressof(sl), C.addressof(sa), C.addressof(sb), 0)
                    P("synchronize")()
                    assert rel_err(D.fetch(sl, len(dl.value)), wl.value) <= tol
                    assert rel_err(D.fetch(su, len(du.value)), wu.value) <= tol
                    n_done += 1
                elif (i, j) in off:
                    c = off[(i, j)]
                    want = copy.deepcopy(c)
                    O.ssssm(nb, want, a, b)
                    sc = D.slot(c)
                    P("ssssm")(nb, C.addressof(sc), C.addressof(sa), C.addressof(sb), 0)
                    P("synchronize")()
                    assert rel_err(D.fetch(sc, len(c.value)), want.value) <= tol, f"ssssm {(i, j, k)}"
                    n_done += 1
        if rows and cols:
            assert n_done > 0
    finally:
        D.close()


@pytest.mark.parametrize("vtype", ["r64", "r32"])
def test_hybrid_batched_same_destination(vtype, ssssm_mode):
    """Several SSSSM tasks with the SAME opdst in one hybrid_batched call (the aggregator's batches,
    src/pangulu_task.c:39-91) must be combined, not raced."""
    nb = 32
    dtype = _lib.VALUE_DTYPES[vtype]
    _lib.load(vtype).pangulu_sm100_set_option(b"ssssm_mode", ssssm_mode)
    blocks = blocks_of(matrices.poisson3d(8), nb, vtype)
    rng = np.random.default_rng(7)
    _randomise(blocks, rng, dtype)
    off, diag = _index(blocks)
    O = _oracle(vtype)
    D = DeviceBlocks(vtype, nb)
    try:
        # destinations with the most updates
        upd = {}
        for k in sorted(diag):
            rows = [i for (i, j) in off if j == k and i > k]
            cols = [j for (i, j) in off if i == k and j > k]
            for i in rows:
                for j in cols:
                    if i == j or (i, j) in off:
                        upd.setdefault((i, j), []).append(k)
        # a batch never reads a block it also updates (operands are finished blocks): keep destinations
        # whose operands are not destinations of the batch and vice versa
        dests, used_as_operand = [], set()
        for d in sorted(upd, key=lambda d: -len(upd[d])):
            ops = {(d[0], k) for k in upd[d]} | {(k, d[1]) for k in upd[d]}
            if d in used_as_operand or ops & set(dests):
                continue
            dests.append(d)
            used_as_operand |= ops
            if len(dests) == 6:
                break
        assert len(upd[dests[0]]) >= 2
        tasks = []
        slots = {}
        want = {}

        def slot_of(key, b):
            if key not in slots:
                slots[key] = D.slot(b)
            return slots[key]

        for (i, j) in dests:
            for k in upd[(i, j)]:
                a, b = off[(i, k)], off[(k, j)]
                sa, sb = slot_of((i, k), a), slot_of((k, j), b)
                if i == j:
                    dl, du = diag[i]
                    if ("d", i) not in want:
                        want[("d", i)] = (copy.deepcopy(dl), copy.deepcopy(du))
                        sl, su = slot_of(("dl", i), dl), slot_of(("du", i), du)
                        D.link(sl, su)
                    O.ssssm_diag(nb, want[("d", i)][0], want[("d", i)][1], a, b)
                    dst = slots[("dl", i)]
                else:
                    c = off[(i, j)]
                    if (i, j) not in want:
                        want[(i, j)] = copy.deepcopy(c)
                    O.ssssm(nb, want[(i, j)], a, b)
                    dst = slot_of((i, j), c)
                t = _lib.Task(row=i, col=j, kernel_id=_lib.TASK_SSSSM, task_level=k, compare_flag=k,
                              opdst=C.addressof(dst), op1=C.addressof(sa), op2=C.addressof(sb))
                tasks.append(t)
        rng.shuffle(tasks)  # order inside a batch is the scheduler's, not ours
        arr = (_lib.Task * len(tasks))(*tasks)
        D.P("hybrid_batched")(nb, len(tasks), C.addressof(arr))
        D.P("synchronize")()
        tol = TOL[vtype]
        for key, w in want.items():
            if key[0] == "d":
                i = key[1]
                assert rel_err(D.fetch(slots[("dl", i)], len(w[0].value)), w[0].value) <= tol
                assert rel_err(D.fetch(slots[("du", i)], len(w[1].value)), w[1].value) <= tol
            else:
                assert rel_err(D.fetch(slots[key], len(w.value)), w.value) <= tol
    finally:
        D.close()


def test_empty_and_padded_blocks():
    """Edge cases: a block whose last rows/columns are padding beyond n, and an SSSSM whose operands have
    empty columns -- must be no-ops there and must not touch memory out of pattern."""
    nb = 16
    A = matrices.laplace2d(5)  # n = 25 -> 2 blocks of 16, the second one padded
    blocks = blocks_of(A, nb, "r64")
    off, diag = _index(blocks)
    O = _oracle("r64")
    D = DeviceBlocks("r64", nb)
    try:
        lo, up = diag[1]
        assert up.colptr[nb] < nb * nb and up.colptr[nb] == up.colptr[9]  # rows >= 9 are padding
        want_lo, want_up = copy.deepcopy(lo), copy.deepcopy(up)
        O.getrf(nb, want_lo, want_up)
        s_lo, s_up = D.slot(lo), D.slot(up)
        D.link(s_lo, s_up)
        D.P("getrf")(nb, C.addressof(s_lo), 0)
        D.P("synchronize")()
        assert rel_err(D.fetch(s_up, len(up.value)), want_up.value) <= 1e-12
        assert rel_err(D.fetch(s_lo, len(lo.value)), want_lo.value) <= 1e-12
    finally:
        D.close()


@pytest.mark.parametrize("vtype", ["r64", "r32", "cr64"])
@pytest.mark.parametrize("getrf_mode", [1, 0, 3], ids=["global-sparse", "smem-staged", "dense-promoted"])
@pytest.mark.parametrize("make,nb", [(lambda: matrices.poisson3d(10), 64), (lambda: matrices.stencil27(7), 200), (lambda: matrices.laplace2d(21), 256)],
                         ids=["poisson3d-10-nb64", "stencil27-7-nb200", "laplace2d-21-nb256-padded"])
def test_getrf_variants(make, nb, getrf_mode, vtype):
    """The three GETRF kernels (global-memory sparse, shared-memory staged, dense-promoted blocked LU) against the
    oracle on every diagonal block of a matrix, values randomised."""
    dtype = _lib.VALUE_DTYPES[vtype]
    L = _lib.load(vtype)
    blocks = blocks_of(make(), nb, vtype)
    rng = np.random.default_rng(99)
    _randomise(blocks, rng, dtype)
    _, diag = _index(blocks)
    O = _oracle(vtype)
    D = DeviceBlocks(vtype, nb)
    L.pangulu_sm100_set_option(b"getrf_mode", getrf_mode)
    try:
        tasks, slots, want = [], [], []
        for k in sorted(diag):
            lo, up = diag[k]
            wl, wu = copy.deepcopy(lo), copy.deepcopy(up)
            O.getrf(nb, wl, wu)
            s_lo, s_up = D.slot(lo), D.slot(up)
            D.link(s_lo, s_up)
            slots.append((s_lo, s_up, len(lo.value), len(up.value)))
            want.append((wl, wu))
            tasks.append(_lib.Task(row=k, col=k, kernel_id=_lib.TASK_GETRF, task_level=k, opdst=C.addressof(s_lo)))
        arr = (_lib.Task * len(tasks))(*tasks)
        D.P("hybrid_batched")(nb, len(tasks), C.addressof(arr))  # all diagonal blocks in ONE batched launch
        D.P("synchronize")()
        tol = TOL[vtype]
        for (s_lo, s_up, nl, nu), (wl, wu) in zip(slots, want):
            assert rel_err(D.fetch(s_lo, nl), wl.value) <= tol
            assert rel_err(D.fetch(s_up, nu), wu.value) <= tol
    finally:
        L.pangulu_sm100_set_option(b"getrf_mode", 0)
        D.close()


@pytest.mark.parametrize("vtype", ["r64", "r32", "cr64", "cr32"])
def test_solve_kernels_against_oracle(vtype):
    """spmv / sptrsv / vecadd of the platform interface (the solve-phase block kernels, reference CPU platform
    pangulu_platform_0100000.c:433-506; the stock CUDA platform leaves them empty, ...0201000.cu:958-979) on device
    slots against the oracle, every off-diagonal block and both halves of every diagonal block."""
    O = _oracle(vtype)
    dtype = np.dtype(_lib.VALUE_DTYPES[vtype])
    nb = 32
    blocks = blocks_of(matrices.poisson3d(8), nb, vtype)
    rng = np.random.default_rng(11)
    _randomise(blocks, rng, dtype)
    D = DeviceBlocks(vtype, nb)
    rnd = lambda n: (rng.uniform(-1, 1, n) + (1j * rng.uniform(-1, 1, n) if dtype.kind == "c" else 0)).astype(dtype)
    tol = TOL[vtype]

    def dev_vec(v):
        p = C.c_void_p()
        D.P("malloc")(C.byref(p), max(v.nbytes, 16))
        D.P("memcpy")(p, v.ctypes.data, v.nbytes, 0)
        return p

    def fetch(p, n):
        out = np.zeros(n, dtype=dtype)
        D.P("memcpy")(out.ctypes.data, p, out.nbytes, 1)
        return out

    try:
        n_m = n_l = n_u = 0
        for b in blocks[:60] + [b for b in blocks if b.brow == b.bcol][:8]:
            slot = D.slot(b)
            x = rnd(nb)
            if b.brow != b.bcol:
                y = rnd(nb)
                want = y.copy()
                O.spmv(nb, b, x, want)
                dx, dy = dev_vec(x), dev_vec(y)
                D.P("spmv")(nb, C.addressof(slot), dx, dy)
                D.P("synchronize")()
                assert rel_err(fetch(dy, nb), want) <= tol
                n_m += 1
            else:
                want = x.copy()
                O.sptrsv(nb, b, want, 1 if b.is_upper else 0)
                dx = dev_vec(x)
                D.P("sptrsv")(nb, C.addressof(slot), dx, 1 if b.is_upper else 0)
                D.P("synchronize")()
                assert rel_err(fetch(dx, nb), want) <= tol
                n_l += not b.is_upper
                n_u += bool(b.is_upper)
        assert n_m and n_l and n_u
        a, c = rnd(1000), rnd(1000)
        da, dc = dev_vec(a), dev_vec(c)
        D.P("vecadd")(1000, da, dc)
        D.P("synchronize")()
        assert np.array_equal(fetch(da, 1000), a + c)
    finally:
        D.close()


@pytest.mark.parametrize("vtype", ["r64", "cr64", "r32"])
def test_gstrs_on_device_matches_host_solve(vtype):
    """pangulu_gstrs with the factors resident in HBM (block substitution on the device, only the right-hand side travels)
    against the same substitution on the host copies of the factors."""
    from pangulu_b200 import PanguLU

    dtype = np.dtype(_lib.VALUE_DTYPES[vtype])
    A = (matrices.circuit_like(900, seed=3, dmax=16, dtype=dtype) if dtype.kind == "c" else matrices.poisson3d(12).astype(dtype))
    n = A.shape[0]
    b = (A @ np.ones(n, dtype=dtype)).astype(dtype)
    xs = []
    for dev in (1, 0):
        s = PanguLU(A, nb=64, mode=1, device_solve=dev)
        try:
            s.gstrf()
            xs.append(s.gstrs(b))
        finally:
            s.finalize()
    tol = 1e-4 if "32" in vtype else 1e-12
    assert np.linalg.norm(xs[0] - xs[1]) / np.linalg.norm(xs[1]) <= tol
    assert np.linalg.norm(A @ xs[0] - b) / np.linalg.norm(b) <= tol
