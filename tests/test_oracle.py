"""Pin the oracle: the C restatement (oracle/pangulu_oracle.c) against the golden vectors generated from the
unmodified reference (tests/golden/*.npz, tests/golden/make_golden.py) and, where oracle/_ref is present, against
live runs of the reference CPU platform -- per kernel and for whole factorisations."""
import copy
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN, TREFETHEN, rel_err, trefethen_matrix
from oracle.oracle import Oracle
from oracle.ref_run import add_csr_mirror, unflatten
from pangulu_b200 import matrices

GOLDEN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "pattern" not in p)


def test_trefethen_fixture_matches_reference_file():
    z = np.load(TREFETHEN)
    A = trefethen_matrix()
    assert np.array_equal(A.indptr, z["indptr"]) and np.array_equal(A.indices, z["indices"]) and np.array_equal(A.data, z["data"])
    ref = "/root/reference/examples/Trefethen_20b.mtx"
    if os.path.exists(ref):
        B = matrices.trefethen(ref)
        assert (abs(A - B)).nnz == 0


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_block_lu_matches_golden_factors(case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    nb = int(z["nb"])
    pre, post = unflatten(z, "pre_", nb), unflatten(z, "post_", nb)
    for b in pre:
        if b.brow > b.bcol:
            add_csr_mirror(b, nb)
    flop = Oracle("r64").block_lu(nb, pre, count_flop=True)
    assert flop == int(z["flop"])  # the reference's own -DPANGULU_PERF counter
    for g, w in zip(pre, post):
        assert np.array_equal(g.rowidx, w.rowidx)
        assert rel_err(g.value, w.value) <= 1e-13


def _live_ref():
    from oracle.ref import ref_available

    if not ref_available("r64"):
        pytest.skip("oracle/_ref not built on this box")


@pytest.mark.parametrize("vtype", ["r64", "r32", "cr64", "cr32"])
def test_oracle_kernels_match_reference_platform_kernels(vtype):
    """Per kernel, on blocks of a 3D Poisson matrix with random values: oracle vs pangulu_platform_0100000_*
    called directly (GETRF/TSTRF/GESSM bit for bit, SSSSM to rounding)."""
    from oracle.ref import RefLib, ref_available

    if not ref_available(vtype):
        pytest.skip("oracle/_ref not built on this box")
    from helpers import blocks_of

    nb = 16
    dtype = np.dtype({"r64": np.float64, "r32": np.float32, "cr64": np.complex128, "cr32": np.complex64}[vtype])
    blocks = blocks_of(matrices.poisson3d(6), nb, vtype)
    rng = np.random.default_rng(3)
    for b in blocks:
        v = rng.uniform(-1, 1, b.value.shape) + (1j * rng.uniform(-1, 1, b.value.shape) if dtype.kind == "c" else 0)
        b.value = v.astype(dtype)
        if b.brow == b.bcol and b.is_upper:
            b.value[b.colptr[:-1][np.diff(b.colptr.astype(np.int64)) > 0]] = 50.0
    off = {(b.brow, b.bcol): b for b in blocks if b.brow != b.bcol}
    lo = next(b for b in blocks if b.brow == b.bcol == 0 and not b.is_upper)
    up = next(b for b in blocks if b.brow == b.bcol == 0 and b.is_upper)
    R, O = RefLib(vtype), Oracle(vtype)
    tol = 1e-5 if vtype in ("r32", "cr32") else 1e-13

    def ref_slot(b):
        return R.make_slot(b.brow, b.bcol, b.is_upper, b.colptr, b.rowidx, b.value, b.rowptr, b.colidx, b.csr2csc)

    rlo, rup = copy.deepcopy(lo), copy.deepcopy(up)
    s_lo, s_up = ref_slot(rlo), ref_slot(rup)
    R.lib.refh_link_diag(s_lo, s_up)
    R.lib.refh_getrf(nb, s_lo)
    olo, oup = copy.deepcopy(lo), copy.deepcopy(up)
    O.getrf(nb, olo, oup)
    assert np.array_equal(olo.value, rlo.value) and np.array_equal(oup.value, rup.value)
    n_t = n_g = n_s = 0
    for (i, j), b in sorted(off.items()):
        if j == 0:
            rb, ob = copy.deepcopy(b), copy.deepcopy(b)
            R.lib.refh_tstrf(nb, ref_slot(rb), s_up)
            O.tstrf(nb, ob, oup)
            assert np.array_equal(ob.value, rb.value)
            n_t += 1
        if i == 0:
            rb, ob = copy.deepcopy(b), copy.deepcopy(b)
            R.lib.refh_gessm(nb, ref_slot(rb), s_lo)
            O.gessm(nb, ob, olo)
            assert np.array_equal(ob.value, rb.value)
            n_g += 1
    for (i, k) in [key for key in off if key[1] == 0]:
        for (k2, j) in [key for key in off if key[0] == 0]:
            if i != j and (i, j) in off:
                a, b2, c = off[(i, 0)], off[(0, j)], off[(i, j)]
                rc, oc = copy.deepcopy(c), copy.deepcopy(c)
                R.lib.refh_ssssm(nb, ref_slot(rc), ref_slot(a), ref_slot(b2))
                O.ssssm(nb, oc, a, b2)
                assert rel_err(oc.value, rc.value) <= tol
                n_s += 1
    assert n_t and n_g and n_s


@pytest.mark.parametrize("vtype", ["r64", "r32", "cr64", "cr32"])
def test_oracle_solve_kernels_match_reference_platform_kernels(vtype):
    """spmv / sptrsv of the solve phase (pangulu_platform_0100000.c:433-506): the oracle's restatement against the reference
    CPU platform called directly, bit for bit (same operations in the same order)."""
    from oracle.ref import RefLib, ref_available

    if not ref_available(vtype):
        pytest.skip("oracle/_ref not built on this box")
    from helpers import blocks_of

    nb = 16
    dtype = np.dtype({"r64": np.float64, "r32": np.float32, "cr64": np.complex128, "cr32": np.complex64}[vtype])
    blocks = blocks_of(matrices.poisson3d(6), nb, vtype)
    rng = np.random.default_rng(5)
    rnd = lambda shape: (rng.uniform(-1, 1, shape) + (1j * rng.uniform(-1, 1, shape) if dtype.kind == "c" else 0)).astype(dtype)
    for b in blocks:
        b.value = rnd(b.value.shape)
        if b.brow == b.bcol and b.is_upper:
            b.value[b.colptr[:-1][np.diff(b.colptr.astype(np.int64)) > 0]] = 4.0
    R, O = RefLib(vtype), Oracle(vtype)
    if not hasattr(R.lib, "refh_spmv"):
        pytest.skip("oracle/_ref predates the solve-kernel harness entries")
    n_m = n_l = n_u = 0
    for b in blocks:
        slot = R.make_slot(b.brow, b.bcol, b.is_upper, b.colptr, b.rowidx, b.value, b.rowptr, b.colidx, b.csr2csc)
        x = rnd(nb)
        if b.brow != b.bcol:
            y0 = rnd(nb)
            yr, yo = y0.copy(), y0.copy()
            R.lib.refh_spmv(nb, slot, x.ctypes.data, yr.ctypes.data)
            O.spmv(nb, b, x, yo)
            assert np.array_equal(yr, yo)
            n_m += 1
        else:
            xr, xo = x.copy(), x.copy()
            R.lib.refh_sptrsv(nb, slot, xr.ctypes.data, 1 if b.is_upper else 0)
            O.sptrsv(nb, b, xo, 1 if b.is_upper else 0)
            assert np.array_equal(xr, xo)
            n_l += not b.is_upper
            n_u += bool(b.is_upper)
    assert n_m and n_l and n_u


def test_oracle_block_lu_matches_live_reference():
    _live_ref()
    from oracle.ref_run import run_reference

    A = matrices.laplace2d(30)
    nb = 32
    z = run_reference(A, nb, perf=True)
    pre, post = unflatten(z, "pre_", nb), unflatten(z, "post_", nb)
    for b in pre:
        if b.brow > b.bcol:
            add_csr_mirror(b, nb)
    flop = Oracle("r64").block_lu(nb, pre, count_flop=True)
    assert flop == int(z["flop"])
    assert max(rel_err(g.value, w.value) for g, w in zip(pre, post)) <= 1e-13
