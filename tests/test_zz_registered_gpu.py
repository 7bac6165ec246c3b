"""The platform REGISTERED in a PanguLU tree the way INTEGRATION.md section 1 describes: id 0201001 / GPU_SM100 in
platform_list.csv and build_list.csv, PANGULU_DEFAULT_PLATFORM pointing at it, the dispatch (pangulu_platform_helper.c,
pangulu_platform_common.h) written by the reference's own build_helper.py, the module compiled with its native id against
PanguLU's own pangulu_common.h (oracle/Makefile: _ref/libpangulu_ref_reg_r64.so, built from a tree of symbolic links by
oracle/overlay/register_platform.py).  The CPU part checks what was generated and linked; the GPU part runs the reference's
pangulu_init / pangulu_gstrf / pangulu_gstrs through it."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, TOL, rel_err
from pangulu_b200 import matrices

LIB = os.path.join(ROOT, "oracle", "_ref", "libpangulu_ref_reg_r64.so")
TREE = os.path.join(ROOT, "oracle", "_ref", "tree")


def test_registered_build_names_the_platform_and_exports_its_symbols():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref registered build not present")
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    from pangulu_b200 import _lib

    for fn in _lib.PLATFORM_FUNCTIONS:
        assert f"pangulu_platform_0201001_{fn}" in exported          # the module under its own id
        assert f"pangulu_platform_{fn}" in exported                   # the reference's generated dispatch
        assert f"pangulu_platform_0201000_{fn}" not in exported       # the stock CUDA platform is not in this build
    helper = os.path.join(TREE, "src", "pangulu_platform_helper.c")
    if os.path.exists(helper):   # (the build tree exists where the reference tree does)
        text = open(helper).read()
        assert "case PANGULU_PLATFORM_GPU_SM100 : pangulu_platform_0201001_getrf(nb, opdst, tid); break;" in text
        common = open(os.path.join(TREE, "src", "platforms", "pangulu_platform_common.h")).read()
        assert "#define PANGULU_PLATFORM_GPU_SM100 0x0201001" in common
        assert '#include "02_NONSHAREDMEM/01_GPU/001_SM100/pangulu_platform_0201001.h"' in common
        assert "PANGULU_PLATFORM_GPU_SM100" in open(os.path.join(TREE, "src", "pangulu_common.h")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,nb", [("poisson3d-12-nb32", lambda: matrices.poisson3d(12), 32), ("stencil27-8-nb128", lambda: matrices.stencil27(8), 128)],
                         ids=["poisson3d-12-nb32", "stencil27-8-nb128"])
def test_reference_runs_on_the_registered_platform(name, make, nb):
    from oracle.ref import ref_available
    from oracle.ref_run import run_reference, unflatten

    if not (ref_available("r64", flavour="reg") and ref_available("r64", flavour="cpu")):
        pytest.skip("oracle/_ref registered / CPU reference builds not present on this box")
    A = make()
    cpu = run_reference(A, nb, flavour="cpu")
    gpu = run_reference(A, nb, flavour="reg")
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
