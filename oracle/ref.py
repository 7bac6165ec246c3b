"""ctypes front-end to the UNMODIFIED PanguLU reference built by oracle/Makefile
into oracle/_ref/ (host code + CPU platform 0100000, see ref_harness.c).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(pangulu_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

VALUE_DTYPES = {"r64": np.float64, "r32": np.float32, "cr64": np.complex128, "cr32": np.complex64}


def ref_lib_path(vtype: str = "r64", perf: bool = False, flavour: str = "cpu") -> str:
    return os.path.join(_REF_DIR, f"libpangulu_ref_{flavour}_{vtype}{'_perf' if perf else ''}.so")


def ref_available(vtype: str = "r64", perf: bool = False, flavour: str = "cpu") -> bool:
    return os.path.exists(ref_lib_path(vtype, perf, flavour))


@dataclass
class RefBlock:
    brow: int
    bcol: int
    is_upper: int
    colptr: np.ndarray  # u32[nb+1]  (row pointer for the upper half of a diagonal block)
    rowidx: np.ndarray  # u16[nnz]
    value: np.ndarray  # T[nnz]
    rowptr: np.ndarray | None = None
    colidx: np.ndarray | None = None
    csr2csc: np.ndarray | None = None


class RefLib:
    """One loaded reference library (one value type, one build flavour)."""

    def __init__(self, vtype: str = "r64", perf: bool = False, flavour: str = "cpu"):
        path = ref_lib_path(vtype, perf, flavour)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.vtype = vtype
        self.dtype = np.dtype(VALUE_DTYPES[vtype])
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        L.refh_init.restype = C.c_void_p
        L.refh_init.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float]
        for name in ("refh_gstrf",):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.refh_gstrs.argtypes = [C.c_void_p, C.c_void_p]
        L.refh_gstrs.restype = None
        L.refh_time.argtypes = [C.c_void_p, C.c_int]
        L.refh_time.restype = C.c_double
        L.refh_flop.restype = C.c_longlong
        for name in ("refh_block_length", "refh_nb", "refh_symbolic_nnz", "refh_num_slots", "refh_num_nondiag"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_longlong
        L.refh_slot.argtypes = [C.c_void_p, C.c_longlong]
        L.refh_slot.restype = C.c_void_p
        L.refh_slot_info.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        L.refh_slot_array.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
        L.refh_slot_array.restype = C.c_void_p
        L.refh_perm.argtypes = [C.c_void_p, C.c_int]
        L.refh_perm.restype = C.c_void_p
        L.refh_make_slot.restype = C.c_void_p
        L.refh_make_slot.argtypes = [C.c_uint32, C.c_uint32, C.c_int] + [C.c_void_p] * 6
        L.refh_link_diag.argtypes = [C.c_void_p, C.c_void_p]
        L.refh_free_slot.argtypes = [C.c_void_p]
        L.refh_getrf.argtypes = [C.c_int, C.c_void_p]
        L.refh_tstrf.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.refh_gessm.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.refh_ssssm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refh_transpose_with_valueidx.argtypes = [C.c_int] + [C.c_void_p] * 5
        if hasattr(L, "refh_spmv"):
            L.refh_spmv.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            L.refh_vecadd.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]
            L.refh_sptrsv.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
        L.refh_slot_offsets.argtypes = [C.c_void_p]
        L.pangulu_oracle_blas_backend.restype = C.c_char_p
        assert L.refh_sizeof_value() == self.dtype.itemsize

    # ---- capabilities --------------------------------------------------
    @property
    def has_metis(self) -> bool:
        return bool(self.lib.refh_has_metis())

    @property
    def has_mc64(self) -> bool:
        return bool(self.lib.refh_has_mc64())

    @property
    def blas_backend(self) -> str:
        return self.lib.pangulu_oracle_blas_backend().decode()

    def struct_offsets(self) -> np.ndarray:
        out = np.zeros(30, dtype=np.int64)
        self.lib.refh_slot_offsets(out.ctypes.data)
        return out

    # ---- end to end ----------------------------------------------------
    def factor(self, n, colptr, rowidx, value, nb, nthread=1, recv_level=0.5) -> "RefFactor":
        colptr = np.ascontiguousarray(colptr, dtype=np.uint64)
        rowidx = np.ascontiguousarray(rowidx, dtype=np.uint32)
        value = np.ascontiguousarray(value, dtype=self.dtype)
        # the reference keeps (and re-allocates through) the arrays it is given: hand it private copies
        # allocated with malloc so its realloc/free calls stay legal
        libc = C.CDLL(None)
        libc.malloc.restype = C.c_void_p
        libc.malloc.argtypes = [C.c_size_t]

        def dup(a):
            p = libc.malloc(max(a.nbytes, 8))
            C.memmove(p, a.ctypes.data, a.nbytes)
            return p

        ctx = self.lib.refh_init(n, int(colptr[-1]), dup(colptr), dup(rowidx), dup(value), nb, nthread, recv_level)
        f = RefFactor(self, ctx, n)
        return f

    # ---- per-kernel oracle ----------------------------------------------
    def make_slot(self, brow, bcol, is_upper, colptr, rowidx, value, rowptr=None, colidx=None, csr2csc=None):
        p = lambda a: None if a is None else a.ctypes.data
        return self.lib.refh_make_slot(brow, bcol, is_upper, p(colptr), p(rowidx), p(value), p(rowptr), p(colidx), p(csr2csc))

    def csr_mirror(self, nb, colptr, rowidx):
        nnz = int(colptr[nb])
        rowptr = np.zeros(nb + 1, dtype=np.uint32)
        colidx = np.zeros(max(nnz, 1), dtype=np.uint16)
        csr2csc = np.zeros(max(nnz, 1), dtype=np.uint32)
        self.lib.refh_transpose_with_valueidx(nb, colptr.ctypes.data, rowidx.ctypes.data, rowptr.ctypes.data,
                                              colidx.ctypes.data, csr2csc.ctypes.data)
        return rowptr, colidx[:nnz], csr2csc[:nnz]


class RefFactor:
    def __init__(self, lib: RefLib, ctx, n):
        self.reflib = lib
        self.L = lib.lib
        self.ctx = ctx
        self.n = n
        self.nb = int(self.L.refh_nb(ctx))
        self.block_length = int(self.L.refh_block_length(ctx))
        self.symbolic_nnz = int(self.L.refh_symbolic_nnz(ctx))

    def gstrf(self) -> float:
        self.L.refh_gstrf(self.ctx)
        return float(self.L.refh_time(self.ctx, 1))

    def gstrs(self, rhs: np.ndarray) -> np.ndarray:
        x = np.array(rhs, dtype=self.reflib.dtype, copy=True)
        self.L.refh_gstrs(self.ctx, x.ctypes.data)
        return x

    def flop(self) -> int:
        return int(self.L.refh_flop())

    def _arr(self, ptr, count, dtype):
        if not ptr or count == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def perms(self):
        n = self.n
        g = lambda w, dt: self._arr(self.L.refh_perm(self.ctx, w), n, dt).copy()
        return {
            "row_perm": g(0, np.uint32), "col_perm": g(1, np.uint32), "metis_perm": g(2, np.uint32),
            "row_scale": g(3, self.reflib.dtype), "col_scale": g(4, self.reflib.dtype),
        }

    def blocks(self, copy=True) -> list[RefBlock]:
        """All owned blocks in the order of bins[0] (src/pangulu_storage.c:239-293):
        off-diagonal blocks in block-CSC order, then (lower, upper) per diagonal block."""
        out = []
        info = np.zeros(5, dtype=np.int64)
        nb = self.nb
        for i in range(int(self.L.refh_num_slots(self.ctx))):
            self.L.refh_slot_info(self.ctx, i, info.ctypes.data)
            brow, bcol, is_upper, nnz, has_csr = (int(v) for v in info)
            a = lambda w, cnt, dt: (self._arr(self.L.refh_slot_array(self.ctx, i, w), cnt, dt).copy() if copy
                                    else self._arr(self.L.refh_slot_array(self.ctx, i, w), cnt, dt))
            b = RefBlock(brow, bcol, is_upper, a(0, nb + 1, np.uint32), a(1, nnz, np.uint16), a(2, nnz, self.reflib.dtype))
            if has_csr:
                b.rowptr, b.colidx, b.csr2csc = a(3, nb + 1, np.uint32), a(4, nnz, np.uint16), a(5, nnz, np.uint32)
            out.append(b)
        return out
