"""Parity at production granularity against the reference itself: 3D Poisson 48^3 (110 592 rows), nb = 256 -- the block
size, the kernels (compact tiles, compact panels, independent groups of diagonal blocks, look-ahead windows, persistent
executor) and the regime (hundreds of block rows) of the benchmark -- factored by the unmodified reference CPU platform
(oracle/_ref, one rank) and by this module on 1 GPU and on 2 / 4 / 8 ranks.

Gates (north star): nonzero structure and permutation bit-exact, L/U values within 1e-12 block-relative, solve residual
no worse than the reference's."""
import os

import numpy as np
import pytest

from helpers import TOL, rel_err
from pangulu_b200 import PanguLU, matrices
from test_multigpu_gpu import run_ranks

pytestmark = pytest.mark.gpu

GRID, NB = 48, 256


@pytest.fixture(scope="module")
def reference_factors(tmp_path_factory):
    from oracle.ref import ref_available

    if not ref_available("r64"):
        pytest.skip("oracle/_ref (the compiled reference) is not on this box")
    from oracle.ref_run import run_reference

    z = run_reference(matrices.poisson3d(GRID), NB, vtype="r64", solve=True)
    path = str(tmp_path_factory.mktemp("ref48") / "poisson48_nb256_reference.npz")
    np.savez(path, **z)
    return path, z


@pytest.mark.parametrize("mode", [1, 2], ids=["resident-plan", "host-buffers"])
def test_one_gpu_matches_reference(reference_factors, mode):
    from oracle.ref_run import unflatten

    _, z = reference_factors
    want = unflatten(z, "post_", NB)
    A = matrices.poisson3d(GRID)
    s = PanguLU(A, nb=NB, mode=mode)
    try:
        s.gstrf()
        if mode == 1:
            s.download()
        got = s.blocks(copy=True)
        assert len(got) == len(want)
        assert np.array_equal(s.perms()["metis_perm"], z["metis_perm"])          # permutation: bit-exact
        worst = 0.0
        for g, w in zip(got, want):
            assert (g.brow, g.bcol, g.is_upper) == (w.brow, w.bcol, w.is_upper)
            assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)   # structure: bit-exact
            worst = max(worst, rel_err(g.value, w.value))
        assert worst <= TOL["r64"], f"L/U values differ from the reference by {worst:.3e}"
        assert s.factor_check() < 1e-12
        b = A @ np.ones(A.shape[0])
        x = s.gstrs(b)
        resid = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        assert resid <= max(4.0 * float(z["resid"]), 1e-15), f"residual {resid:.3e} vs reference {float(z['resid']):.3e}"
    finally:
        s.finalize()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_matches_reference(reference_factors, world):
    path, _ = reference_factors
    run_ranks(world, {"GRID": str(GRID), "NB": str(NB), "EXPECT": path}, timeout=900)


def test_reference_gstrf_plan_path_matches_reference(reference_factors):
    """Production granularity through the reference's OWN pangulu_init / pangulu_gstrf / pangulu_gstrs with the integration
    overlay (tests/test_dropin_gpu.py): PanguLU's storage, this module's plan path underneath."""
    from oracle.ref import ref_available
    from oracle.ref_run import run_reference, unflatten

    if not ref_available("r64", flavour="plan"):
        pytest.skip("oracle/_ref plan build not present on this box")
    _, z = reference_factors
    gpu = run_reference(matrices.poisson3d(GRID), NB, flavour="plan")
    assert np.array_equal(z["metis_perm"], gpu["metis_perm"])
    want, got = unflatten(z, "post_", NB), unflatten(gpu, "post_", NB)
    assert len(want) == len(got)
    worst = 0.0
    for g, w in zip(got, want):
        assert (g.brow, g.bcol, g.is_upper) == (w.brow, w.bcol, w.is_upper)
        assert np.array_equal(g.colptr, w.colptr) and np.array_equal(g.rowidx, w.rowidx)
        worst = max(worst, rel_err(g.value, w.value))
    assert worst <= TOL["r64"], worst
    assert float(gpu["resid"]) <= max(4 * float(z["resid"]), 1e-15)
    print("plan path behind the reference's pangulu_gstrf, 48^3 nb 256: schedule %.2f s, plan %.2f s, device %.3f s, download %.3f s; "
          "reference CPU gstrf %.2f s" % (*[float(x) for x in gpu["plan_times"]], float(z["t_gstrf"])))
