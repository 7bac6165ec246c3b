"""Drop-in check on the GPU: the UNMODIFIED reference host code (pangulu_init / pangulu_gstrf / pangulu_gstrs, its
dynamic scheduler, block storage, kernel interface and generated platform dispatch, compiled with -DGPU_OPEN) running on
top of this module, whose object file takes the place of the stock CUDA platform (oracle/_ref/libpangulu_ref_gpu_*.so,
built by oracle/Makefile with -DPANGULU_SM100_ID=0201000).  Expected: same structure, same permutation, L/U values
within tolerance of the reference's own CPU platform, residual no worse."""
import numpy as np
import pytest

from helpers import TOL, rel_err, trefethen_matrix
from pangulu_b200 import matrices

pytestmark = pytest.mark.gpu

CASES = [
    ("trefethen-nb4", lambda: trefethen_matrix(), 4),
    ("trefethen-nb10", lambda: trefethen_matrix(), 10),
    ("poisson3d-12-nb32", lambda: matrices.poisson3d(12), 32),
    ("laplace2d-48-nb64", lambda: matrices.laplace2d(48), 64),
    ("stencil27-8-nb128", lambda: matrices.stencil27(8), 128),
]


@pytest.mark.parametrize("name,make,nb", CASES, ids=[c[0] for c in CASES])
def test_reference_scheduler_on_sm100_platform(name, make, nb):
    from oracle.ref import ref_available
    from oracle.ref_run import run_reference, unflatten

    if not (ref_available("r64", flavour="gpu") and ref_available("r64", flavour="cpu")):
        pytest.skip("oracle/_ref GPU/CPU reference builds not present on this box")
    A = make()
    cpu = run_reference(A, nb, flavour="cpu")
    gpu = run_reference(A, nb, flavour="gpu")
    assert np.array_equal(cpu["metis_perm"], gpu["metis_perm"])
    want, got = unflatten(cpu, "post_", nb), unflatten(gpu, "post_", nb)
    assert len(want) == len(got)
    worst = 0.0
    for g, w in zip(got, want):
        assert (g.brow, g.bcol, g.is_upper) == (w.brow, w.bcol, w.is_upper)
        assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)
        assert np.isfinite(g.value).all()
        worst = max(worst, rel_err(g.value, w.value))
    assert worst <= TOL["r64"], worst
    assert float(gpu["resid"]) <= max(4 * float(cpu["resid"]), 1e-15)


PLAN_CASES = CASES + [("poisson3d-20-nb64", lambda: matrices.poisson3d(20), 64)]


@pytest.mark.parametrize("name,make,nb", PLAN_CASES, ids=[c[0] for c in PLAN_CASES])
def test_reference_gstrf_runs_the_plan_path(name, make, nb):
    """The integration overlay (oracle/overlay/apply_overlay.py, 12 lines inside PanguLU's pangulu_gstrf): the reference's
    own pangulu_init (its reordering, symbolic factorisation and block storage) followed by its pangulu_gstrf, which hands
    the block storage to this module's static plan -- compact tile / compact panel kernels, independent groups of
    diagonal blocks, persistent executor -- and the reference's own pangulu_gstrs on the returned factors."""
    from oracle.ref import ref_available
    from oracle.ref_run import run_reference, unflatten

    if not (ref_available("r64", flavour="plan") and ref_available("r64", flavour="cpu")):
        pytest.skip("oracle/_ref plan/CPU reference builds not present on this box")
    A = make()
    cpu = run_reference(A, nb, flavour="cpu")
    gpu = run_reference(A, nb, flavour="plan")
    assert np.array_equal(cpu["metis_perm"], gpu["metis_perm"])
    want, got = unflatten(cpu, "post_", nb), unflatten(gpu, "post_", nb)
    assert len(want) == len(got)
    worst = 0.0
    for g, w in zip(got, want):
        assert (g.brow, g.bcol, g.is_upper) == (w.brow, w.bcol, w.is_upper)
        assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)
        assert np.isfinite(g.value).all()
        worst = max(worst, rel_err(g.value, w.value))
    assert worst <= TOL["r64"], worst
    assert float(gpu["resid"]) <= max(4 * float(cpu["resid"]), 1e-15)
    assert gpu["plan_times"][2] > 0, "pangulu_gstrf did not take the plan path"


def test_mc64_build_of_the_reference_on_the_plan_path():
    """BASELINE's "MC64+METIS" configuration: the reference compiled with its own -DPANGULU_MC64 (scaling + row matching in
    pangulu_init, src/pangulu_reordering.c:149-680) on the CPU platform and on this module's plan path; the input is a
    row- and column-scaled Poisson matrix, so the scaling vectors are not trivial.  Scaling, permutations and structure
    bit-exact (the same host code computes them), values and residual as everywhere."""
    import scipy.sparse as sp

    from oracle.ref import ref_available
    from oracle.ref_run import run_reference, unflatten

    if not (ref_available("r64", flavour="planmc64") and ref_available("r64", flavour="cpumc64")):
        pytest.skip("oracle/_ref MC64 builds not present on this box")
    A0 = matrices.poisson3d(16)
    n = A0.shape[0]
    rng = np.random.default_rng(3)
    A = sp.csc_matrix(sp.diags(10.0 ** rng.uniform(-1, 1, n)) @ A0 @ sp.diags(10.0 ** rng.uniform(-1, 1, n)))
    A.sort_indices()
    nb = 64
    cpu = run_reference(A, nb, flavour="cpumc64")
    gpu = run_reference(A, nb, flavour="planmc64")
    assert not np.allclose(cpu["row_scale"], 1.0), "MC64 scaling did not run"
    for k in ("row_perm", "col_perm", "metis_perm", "row_scale", "col_scale"):
        assert np.array_equal(cpu[k], gpu[k]), k
    want, got = unflatten(cpu, "post_", nb), unflatten(gpu, "post_", nb)
    assert len(want) == len(got)
    worst = 0.0
    for g, w in zip(got, want):
        assert (g.brow, g.bcol, g.is_upper) == (w.brow, w.bcol, w.is_upper)
        assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)
        worst = max(worst, rel_err(g.value, w.value))
    assert worst <= TOL["r64"], worst
    assert float(gpu["resid"]) <= max(4 * float(cpu["resid"]), 1e-15)
    assert gpu["plan_times"][2] > 0, "pangulu_gstrf did not take the plan path"
