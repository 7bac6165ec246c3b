"""ctypes front-end to the C restatement of PanguLU's block kernels (oracle/pangulu_oracle.c) and a
block right-looking LU built from them.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product package pangulu_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VALUE_DTYPES = {"r64": np.float64, "r32": np.float32, "cr64": np.complex128, "cr32": np.complex64}
DTYPE_TO_VTYPE = {np.dtype(v): k for k, v in VALUE_DTYPES.items()}


def build(vtypes=("r64", "r32", "cr64", "cr32"), with_reference: bool = True) -> None:
    """Compile the oracle (and, where /root/reference exists, the reference build under oracle/_ref)."""
    targets = ["oracle"] + (["ref"] if with_reference else [])
    subprocess.run(["make", "-C", _HERE, "-j4", f"TYPES={' '.join(vtypes)}"] + targets, check=True,
                   stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self, vtype: str = "r64"):
        path = os.path.join(_HERE, f"liboracle_{vtype}.so")
        if not os.path.exists(path):
            build((vtype,), with_reference=False)
        self.vtype = vtype
        self.dtype = np.dtype(VALUE_DTYPES[vtype])
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        for name, nargs in (("oracle_getrf", 6), ("oracle_tstrf", 7), ("oracle_gessm", 6), ("oracle_ssssm", 9),
                            ("oracle_ssssm_diag", 12), ("oracle_spmv", 5)):
            f = getattr(L, name)
            f.argtypes = [C.c_int] + [vp] * nargs
            f.restype = None
        L.oracle_sptrsv.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int]
        for name, nargs in (("oracle_getrf_flop", 4), ("oracle_tstrf_flop", 4), ("oracle_gessm_flop", 4), ("oracle_ssssm_flop", 3)):
            f = getattr(L, name)
            f.argtypes = [C.c_int] + [vp] * nargs
            f.restype = C.c_longlong
        assert L.oracle_sizeof_value() == self.dtype.itemsize

    @staticmethod
    def _p(a):
        return a.ctypes.data

    # blocks are objects with .colptr .rowidx .value (.rowptr .colidx .csr2csc for strictly-lower blocks)
    def getrf(self, nb, lower, upper):
        p = self._p
        self.lib.oracle_getrf(nb, p(lower.colptr), p(lower.rowidx), p(lower.value), p(upper.colptr), p(upper.rowidx), p(upper.value))

    def tstrf(self, nb, b, upper):
        p = self._p
        self.lib.oracle_tstrf(nb, p(b.rowptr), p(b.colidx), p(b.csr2csc), p(b.value), p(upper.colptr), p(upper.rowidx), p(upper.value))

    def gessm(self, nb, b, lower):
        p = self._p
        self.lib.oracle_gessm(nb, p(b.colptr), p(b.rowidx), p(b.value), p(lower.colptr), p(lower.rowidx), p(lower.value))

    def ssssm(self, nb, c, a, b):
        p = self._p
        self.lib.oracle_ssssm(nb, p(c.colptr), p(c.rowidx), p(c.value), p(a.colptr), p(a.rowidx), p(a.value),
                              p(b.colptr), p(b.rowidx), p(b.value))

    def ssssm_diag(self, nb, lower, upper, a, b):
        p = self._p
        self.lib.oracle_ssssm_diag(nb, p(lower.colptr), p(lower.rowidx), p(lower.value), p(upper.colptr), p(upper.rowidx),
                                   p(upper.value), p(a.colptr), p(a.rowidx), p(a.value), p(b.colptr), p(b.rowidx), p(b.value))

    # solve-phase block kernels (pangulu_platform_0100000.c:433-506): y -= A x;  in-place triangular solve of a diagonal half
    def spmv(self, nb, a, x, y):
        p = self._p
        self.lib.oracle_spmv(nb, p(a.colptr), p(a.rowidx), p(a.value), p(x), p(y))

    def sptrsv(self, nb, s, x, uplo):
        p = self._p
        self.lib.oracle_sptrsv(nb, p(s.colptr), p(s.rowidx), p(s.value), p(x), int(uplo))

    def flop_getrf(self, nb, lower, upper):
        p = self._p
        return self.lib.oracle_getrf_flop(nb, p(lower.colptr), p(lower.rowidx), p(upper.colptr), p(upper.rowidx))

    def flop_tstrf(self, nb, b, upper):
        p = self._p
        return self.lib.oracle_tstrf_flop(nb, p(b.rowptr), p(b.colidx), p(upper.colptr), p(upper.rowidx))

    def flop_gessm(self, nb, b, lower):
        p = self._p
        return self.lib.oracle_gessm_flop(nb, p(b.colptr), p(b.rowidx), p(lower.colptr), p(lower.rowidx))

    def flop_ssssm(self, nb, a, b):
        p = self._p
        return self.lib.oracle_ssssm_flop(nb, p(a.colptr), p(b.colptr), p(b.rowidx))

    # ---- block right-looking LU with the oracle kernels (restates the task graph of
    # src/pangulu_numeric.c: GETRF(k) -> TSTRF(i,k), GESSM(k,j) -> SSSSM(i,j,k), level by level) ----
    def block_lu(self, nb: int, blocks: list, count_flop: bool = False) -> int:
        """Factor in place. `blocks`: objects with brow, bcol, is_upper and the arrays above (the list the
        driver / the reference harness returns: off-diagonal blocks, then (lower, upper) per diagonal block)."""
        off = {}
        diag = {}
        for b in blocks:
            if b.brow == b.bcol:
                diag.setdefault(b.brow, [None, None])[1 if b.is_upper else 0] = b
            else:
                off[(b.brow, b.bcol)] = b
        col_rows, row_cols = {}, {}
        for (i, j) in off:
            if i > j:
                col_rows.setdefault(j, []).append(i)
            else:
                row_cols.setdefault(i, []).append(j)
        flop = 0
        for k in sorted(diag):
            lo, up = diag[k]
            if count_flop:
                flop += self.flop_getrf(nb, lo, up)
            self.getrf(nb, lo, up)
            rows = sorted(col_rows.get(k, []))
            cols = sorted(row_cols.get(k, []))
            for i in rows:
                if count_flop:
                    flop += self.flop_tstrf(nb, off[(i, k)], up)
                self.tstrf(nb, off[(i, k)], up)
            for j in cols:
                if count_flop:
                    flop += self.flop_gessm(nb, off[(k, j)], lo)
                self.gessm(nb, off[(k, j)], lo)
            for i in rows:
                for j in cols:
                    a, b = off[(i, k)], off[(k, j)]
                    if i == j:
                        dl, du = diag[i]
                        self.ssssm_diag(nb, dl, du, a, b)
                    elif (i, j) in off:
                        self.ssssm(nb, off[(i, j)], a, b)
                    else:
                        continue
                    if count_flop:
                        flop += self.flop_ssssm(nb, a, b)
        return flop
