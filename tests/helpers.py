"""Shared helpers of the parity tests: matrices, block extraction, device slots built through the C ABI."""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace

import numpy as np

from pangulu_b200 import _lib, matrices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TREFETHEN = os.path.join(GOLDEN, "Trefethen_20b_pattern.npz")

# value tolerances of the north star: L/U values within 1e-12 relative (R64/CR64) or 1e-5 (R32/CR32)
TOL = {"r64": 1e-12, "cr64": 1e-12, "r32": 1e-5, "cr32": 1e-5}


def trefethen_matrix(dtype=np.float64):
    """The reference's shipped fixture examples/Trefethen_20b.mtx, regenerated from its definition
    (the 19x19 trailing minor of the 20x20 Trefethen matrix: primes 3..71 on the diagonal, ones at |i-j| in
    {1,2,4,8,16}); tests/test_oracle.py checks it against the .mtx file where the reference tree is present."""
    import scipy.sparse as sp

    n = 19
    primes = [3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71]
    A = sp.lil_matrix((n, n), dtype=dtype)
    for i in range(n):
        A[i, i] = primes[i]
        for d in (1, 2, 4, 8, 16):
            if i + d < n:
                A[i, i + d] = 1
                A[i + d, i] = 1
    A = sp.csc_matrix(A)
    A.sort_indices()
    return A


def rel_err(got: np.ndarray, want: np.ndarray) -> float:
    """max |got-want| / max|want| over one block (blockwise relative error)."""
    if want.size == 0:
        return 0.0
    scale = float(np.abs(want).max())
    return float(np.abs(got - want).max() / (scale if scale > 0 else 1.0))


class DeviceBlocks:
    """Device mirrors of host blocks, allocated and filled through the platform C ABI
    (<id>_malloc / <id>_memcpy), with pangulu_storage_slot_t structs wired like the reference's storage."""

    def __init__(self, vtype: str, nb: int):
        self.L = _lib.load(vtype)
        self.P = self.L.platform
        self.nb = nb
        self.dtype = np.dtype(_lib.VALUE_DTYPES[vtype])
        self._keep = []
        self._dev = []

    def _to_dev(self, arr: np.ndarray) -> int:
        p = C.c_void_p()
        self.P("malloc")(C.byref(p), max(arr.nbytes, 16))
        if arr.nbytes:
            self.P("memcpy")(p, arr.ctypes.data, arr.nbytes, 0)
        self._dev.append(p)
        return p.value

    def slot(self, b) -> _lib.Slot:
        """b: object with brow, bcol, is_upper, colptr, rowidx, value (+ rowptr/colidx/csr2csc)."""
        s = _lib.Slot()
        s.brow_pos, s.bcol_pos, s.is_upper = b.brow, b.bcol, b.is_upper
        colptr = np.ascontiguousarray(b.colptr, dtype=np.uint32)
        rowidx = np.ascontiguousarray(b.rowidx, dtype=np.uint16)
        value = np.ascontiguousarray(b.value, dtype=self.dtype)
        b.colptr, b.rowidx, b.value = colptr, rowidx, value
        self._keep += [colptr, rowidx, value]
        s.columnpointer, s.rowindex, s.value = colptr.ctypes.data, rowidx.ctypes.data, value.ctypes.data
        s.d_value = self._to_dev(value)
        diag_upper = b.brow == b.bcol and b.is_upper
        if diag_upper:
            s.d_rowpointer, s.d_columnindex = self._to_dev(colptr), self._to_dev(rowidx)
        else:
            s.d_columnpointer, s.d_rowindex = self._to_dev(colptr), self._to_dev(rowidx)
        if b.brow > b.bcol:
            rp = np.ascontiguousarray(b.rowptr, dtype=np.uint32)
            ci = np.ascontiguousarray(b.colidx, dtype=np.uint16)
            mp = np.ascontiguousarray(b.csr2csc, dtype=np.uint32)
            self._keep += [rp, ci, mp]
            s.rowpointer, s.columnindex, s.idx_of_csc_value_for_csr = rp.ctypes.data, ci.ctypes.data, mp.ctypes.data
            s.d_rowpointer, s.d_columnindex, s.d_idx_of_csc_value_for_csr = self._to_dev(rp), self._to_dev(ci), self._to_dev(mp)
        self._keep.append(s)
        return s

    @staticmethod
    def link(lower: _lib.Slot, upper: _lib.Slot):
        lower.related_block = C.addressof(upper)
        upper.related_block = C.addressof(lower)

    def fetch(self, s: _lib.Slot, nnz: int) -> np.ndarray:
        out = np.zeros(nnz, dtype=self.dtype)
        if nnz:
            self.P("memcpy")(out.ctypes.data, s.d_value, out.nbytes, 1)
        return out

    def close(self):
        for p in self._dev:
            self.P("free")(p)
        self._dev = []


def blocks_of(A, nb, vtype="r64"):
    """Blocks (pattern + initial values) of A from the driver's host-only structure mode, as independent
    copies with CSR mirrors -- inputs for per-kernel tests."""
    from pangulu_b200 import PanguLU

    s = PanguLU(A.astype(_lib.VALUE_DTYPES[vtype]), nb=nb, host_only=True)
    out = [SimpleNamespace(**vars(b)) for b in s.blocks(copy=True)]
    s.finalize()
    return out
