"""Whole-factorisation parity on the GPU through the reference-shaped API
(pangulu_init -> pangulu_gstrf -> pangulu_gstrs of include/pangulu_sm100_driver.h).

Gates (north star): nonzero structure and permutation indices bit-exact, L/U values within 1e-12
(R64/CR64) / 1e-5 (R32/CR32) block-wise relative, solve residual no worse than the reference's.
The expected factors come from the unmodified reference run in a subprocess (oracle/_ref, prebuilt) when it
is present, otherwise from the C oracle's block LU (oracle/pangulu_oracle.c).
"""
import copy
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import TOL, rel_err, trefethen_matrix
from pangulu_b200 import PanguLU, _lib, matrices

pytestmark = pytest.mark.gpu

CASES = [
    ("trefethen-nb4", lambda: trefethen_matrix(), 4),
    ("trefethen-nb10", lambda: trefethen_matrix(), 10),
    ("trefethen-nb19", lambda: trefethen_matrix(), 19),
    ("laplace2d-40-nb32", lambda: matrices.laplace2d(40), 32),
    ("poisson3d-12-nb64", lambda: matrices.poisson3d(12), 64),
    ("stencil27-8-nb64", lambda: matrices.stencil27(8), 64),
    ("poisson3d-16-nb256", lambda: matrices.poisson3d(16), 256),
]


def _expected(A, nb, vtype):
    """(blocks after gstrf, metis_perm or None, residual or None) from the reference, else the C oracle."""
    from oracle.ref import ref_available

    if ref_available(vtype):
        from oracle.ref_run import run_reference, unflatten

        z = run_reference(A, nb, vtype=vtype)
        return unflatten(z, "post_", nb), z["metis_perm"], float(z["resid"])
    from oracle.oracle import Oracle

    s = PanguLU(A, nb=nb, host_only=True)
    blocks = [SimpleNamespace(**vars(b)) for b in s.blocks(copy=True)]
    s.finalize()
    Oracle(vtype).block_lu(nb, blocks)
    return blocks, None, None


@pytest.mark.parametrize("mode", [0, 1], ids=["abi-host-buffers", "resident-plan"])
@pytest.mark.parametrize("name,make,nb", CASES, ids=[c[0] for c in CASES])
def test_gstrf_r64(name, make, nb, mode):
    A = make().astype(np.float64)
    n = A.shape[0]
    want, metis_perm, ref_resid = _expected(A, nb, "r64")
    s = PanguLU(A, nb=nb, mode=mode)
    try:
        s.gstrf()
        if mode == 1:
            s.download()
        got = s.blocks(copy=True)
        assert len(got) == len(want)
        if metis_perm is not None:
            assert np.array_equal(s.perms()["metis_perm"], metis_perm)
        worst = 0.0
        for g, w in zip(got, want):
            assert (g.brow, g.bcol, g.is_upper) == (w.brow, w.bcol, w.is_upper)
            assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)  # structure: bit-exact
            worst = max(worst, rel_err(g.value, w.value))
        assert worst <= TOL["r64"], f"L/U values differ by {worst:.3e}"
        assert s.factor_check() < 1e-12
        b = A @ np.ones(n)
        x = s.gstrs(b)
        resid = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        bound = 1e-13 if ref_resid is None else max(4.0 * ref_resid, 1e-15)
        assert resid <= bound, f"residual {resid:.3e} vs reference {ref_resid}"
    finally:
        s.finalize()


@pytest.mark.parametrize("vtype", ["r32", "cr64", "cr32"])
def test_gstrf_other_types(vtype):
    dtype = _lib.VALUE_DTYPES[vtype]
    nb = 32
    if np.dtype(dtype).kind == "c":
        A = matrices.circuit_like(600, seed=11, dmax=20, dtype=dtype)
    else:
        A = matrices.poisson3d(10).astype(dtype)
    n = A.shape[0]
    want, metis_perm, ref_resid = _expected(A, nb, vtype)
    s = PanguLU(A, nb=nb, mode=0)
    try:
        s.gstrf()
        got = s.blocks(copy=True)
        assert len(got) == len(want)
        worst = 0.0
        for g, w in zip(got, want):
            assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)
            worst = max(worst, rel_err(g.value, w.value))
        assert worst <= TOL[vtype], f"L/U values differ by {worst:.3e}"
        b = (A @ np.ones(n, dtype=dtype)).astype(dtype)
        x = s.gstrs(b)
        resid = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        assert resid <= (1e-4 if "32" in vtype else 1e-12)
    finally:
        s.finalize()


@pytest.mark.parametrize("ssssm_mode", [1, 2], ids=["all-sparse", "all-tile"])
def test_gstrf_forced_ssssm_path(ssssm_mode):
    """The two SSSSM kernels must be interchangeable: force every update through one of them."""
    A = matrices.poisson3d(12)
    nb = 64
    want, _, _ = _expected(A, nb, "r64")
    L = _lib.load("r64")
    L.pangulu_sm100_set_option(b"ssssm_mode", ssssm_mode)
    s = PanguLU(A, nb=nb, mode=0)
    try:
        s.gstrf()
        worst = max(rel_err(g.value, w.value) for g, w in zip(s.blocks(copy=True), want))
        assert worst <= TOL["r64"]
    finally:
        L.pangulu_sm100_set_option(b"ssssm_mode", 0)
        s.finalize()


def test_refactor_after_reset():
    """gstrf twice on one handle (values reset in between) gives the same factors: idempotence of the
    reset + factor cycle used by bench.py."""
    A = matrices.poisson3d(10)
    s = PanguLU(A, nb=32, mode=1)
    try:
        s.gstrf()
        s.download()
        first = [b.value.copy() for b in s.blocks(copy=True)]
        s.reset_values(True)
        s.gstrf()
        s.download()
        second = [b.value for b in s.blocks(copy=True)]
        for a, b in zip(first, second):
            assert rel_err(a, b) <= 1e-13
    finally:
        s.finalize()


@pytest.mark.parametrize("executor", [1, 0], ids=["persistent", "per-wave-launches"])
@pytest.mark.parametrize("vtype,threshold", [("r64", 0), ("r64", 10), ("r32", 0), ("cr64", 0), ("cr32", 0)])
def test_gstrf_compact_dense_path(vtype, threshold, executor):
    """SSSSM on row/column-compacted dense operand copies (sm100_compact.cuh): with threshold 0 EVERY update of the
    resident plan goes through the compact tile kernel (DMMA for R64, FMA tiles otherwise), and the panel solves
    maintain the compact copies.  Same gates as the other paths."""
    dtype = _lib.VALUE_DTYPES[vtype]
    nb = 64 if vtype == "r64" else 32
    if np.dtype(dtype).kind == "c":
        A = matrices.circuit_like(700, seed=5, dmax=20, dtype=dtype)
    else:
        A = (matrices.stencil27(9) if threshold == 0 else matrices.poisson3d(14)).astype(dtype)
    n = A.shape[0]
    want, _, _ = _expected(A, nb, vtype)
    L = _lib.load(vtype)
    L.pangulu_sm100_set_option(b"compact", 1)
    L.pangulu_sm100_set_option(b"compact_threshold", threshold)
    L.pangulu_sm100_set_option(b"compact_block_threshold", 0 if threshold == 0 else 1)
    L.pangulu_sm100_set_option(b"executor", executor)
    s = PanguLU(A, nb=nb, mode=1)
    try:
        L.pangulu_sm100_stats_reset()
        s.gstrf()
        s.download()
        if executor == 0:
            assert L.pangulu_sm100_stats_launches(7) > 0, "compact tile kernel did not run"
        tol = TOL[vtype]
        worst = max(rel_err(g.value, w.value) for g, w in zip(s.blocks(copy=True), want))
        assert worst <= tol, f"L/U values differ by {worst:.3e}"
        s.reset_values(True)   # the compact copies are rewritten by the second factorisation
        s.gstrf()
        s.download()
        worst = max(rel_err(g.value, w.value) for g, w in zip(s.blocks(copy=True), want))
        assert worst <= tol, f"second run: L/U values differ by {worst:.3e}"
        b = (A @ np.ones(n, dtype=dtype)).astype(dtype)
        x = s.gstrs(b)
        resid = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        assert resid <= (1e-4 if "32" in vtype else 1e-12)
    finally:
        L.pangulu_sm100_set_option(b"compact_threshold", 0)      # (the defaults)
        L.pangulu_sm100_set_option(b"compact_block_threshold", 0)
        L.pangulu_sm100_set_option(b"executor", 1)
        s.finalize()


def test_factor_file_resolve_without_refactoring(tmp_path):
    """pangulu_init + load_factors + pangulu_gstrs: a second handle of the same matrix solves with the factors a first
    handle wrote to a file (PanguLU's packed block format), without running gstrf."""
    A = matrices.poisson3d(12)
    n = A.shape[0]
    b = A @ np.ones(n)
    s = PanguLU(A, nb=64, mode=1)
    path = tmp_path / "factors.pglu"
    try:
        s.gstrf()
        s.download()
        x0 = s.gstrs(b)
        s.save_factors(path)
    finally:
        s.finalize()
    t = PanguLU(A, nb=64, mode=1)
    try:
        assert t.load_factors(path) == int(t.info("num_slots"))
        x1 = t.gstrs(b)
    finally:
        t.finalize()
    assert np.array_equal(x0, x1) or np.linalg.norm(x0 - x1) / np.linalg.norm(x0) < 1e-13
    assert np.linalg.norm(A @ x1 - b) / np.linalg.norm(b) < 1e-12


@pytest.mark.parametrize("vtype", ["r64", "cr64"])
def test_device_preprocessing_value_scatter_and_csr_mirrors(vtype):
    """Preprocessing on the GPU.  (1) The CSR mirrors of the strictly-lower blocks rebuilt by a kernel are
    bit-identical to the host-built ones (reference: pangulu_transpose_struct_with_valueidx_inblock).  (2) New values on
    the same pattern scattered into the block value region by a kernel give the same factors as a solver built from
    the new matrix (reference: the host value scatter of src/pangulu_conversion.c:241-350)."""
    dtype = _lib.VALUE_DTYPES[vtype]
    if np.dtype(dtype).kind == "c":
        A = matrices.circuit_like(900, seed=7, dmax=24, dtype=dtype)
    else:
        A = matrices.poisson3d(13).astype(dtype)
    nb = 64
    rng = np.random.default_rng(5)
    A2 = A.copy()
    A2.data = A.data * (1.0 + 0.05 * rng.uniform(-1, 1, A.nnz))   # same pattern, still diagonally dominant
    want = PanguLU(A2, nb=nb, mode=1)
    s = PanguLU(A, nb=nb, mode=1)
    try:
        assert s.device_mirrors(clear_first=True) == 0, "device-built CSR mirrors differ from the host-built ones"
        want.gstrf(); want.download()
        s.gstrf()                                   # factor A first: the update must also clear the old factors
        assert s.update_values(A2.data) > 0
        s.gstrf(); s.download()
        worst = max(rel_err(g.value, w.value) for g, w in zip(s.blocks(copy=True), want.blocks(copy=True)))
        assert worst <= TOL[vtype], f"factors after the device scatter differ by {worst:.3e}"
        assert s.factor_check() < 1e-12             # (against the NEW matrix: the host copy of the reordered matrix followed)
        b = (A2 @ np.ones(A.shape[0], dtype=dtype)).astype(dtype)
        x = s.gstrs(b)
        assert float(np.linalg.norm(A2 @ x - b) / np.linalg.norm(b)) <= 1e-12
    finally:
        s.finalize(); want.finalize()
