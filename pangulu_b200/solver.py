"""Python face of the host driver: the reference's five-call API (include/pangulu.h:11-15) as a class.

    s = PanguLU(A_csc, nb=256)      # pangulu_init  (reorder, symbolic, blocks, schedule, device mirror)
    s.gstrf()                        # pangulu_gstrf (numeric LU on the GPU through the platform C ABI)
    x = s.gstrs(b)                   # pangulu_gstrs
    s.finalize()                     # pangulu_finalize

All arithmetic of the factorisation happens in the CUDA extension; this file only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

DTYPE_TO_VTYPE = {np.dtype(np.float64): "r64", np.dtype(np.float32): "r32",
                  np.dtype(np.complex128): "cr64", np.dtype(np.complex64): "cr32"}


@dataclass
class Block:
    """Host view of one block slot (arrays alias the driver's pinned buffer unless copied)."""
    brow: int
    bcol: int
    is_upper: int
    owned: bool
    colptr: np.ndarray
    rowidx: np.ndarray
    value: np.ndarray
    rowptr: np.ndarray | None = None
    colidx: np.ndarray | None = None
    csr2csc: np.ndarray | None = None


def _as_array(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


def driver_config(vtype: str = "r64", **kw) -> None:
    """Set driver options (include/pangulu_sm100_driver.h) before constructing a solver."""
    L = _lib.load(vtype)
    for k, v in kw.items():
        if L.pangulu_sm100_driver_config(k.encode(), int(v)) != 0:
            raise KeyError(k)


class PanguLU:
    def __init__(self, A, nb: int = 256, *, ordering: int = -1, mode: int = 0, rank: int = 0, nproc: int = 1,
                 host_only: bool = False, nthread: int = 1, verbose: bool = False, schedule: int = 2, lookahead: int = 2,
                 split_diag: int = 1, device_solve: int = 1):
        """A: scipy.sparse matrix (any format; converted to sorted CSC) of dtype f64/f32/c128/c64."""
        import scipy.sparse as sp

        A = sp.csc_matrix(A)
        A.sort_indices()
        self.dtype = np.dtype(A.dtype)
        if self.dtype not in DTYPE_TO_VTYPE:
            raise TypeError(f"unsupported value type {self.dtype}")
        self.vtype = DTYPE_TO_VTYPE[self.dtype]
        self.L = _lib.load(self.vtype)
        self.n = A.shape[0]
        self._host_only = bool(host_only)
        driver_config(self.vtype, ordering=ordering, mode=mode, rank=rank, nproc=nproc, host_only=int(host_only),
                      verbose=int(verbose), schedule=schedule, lookahead=lookahead, split_diag=split_diag, device_solve=device_solve)
        self._colptr = np.ascontiguousarray(A.indptr, dtype=np.uint64)
        self._rowidx = np.ascontiguousarray(A.indices, dtype=np.uint32)
        self._value = np.ascontiguousarray(A.data, dtype=self.dtype)
        opt = _lib.InitOptions(nthread=nthread, nb=nb, gpu_kernel_warp_per_block=4, gpu_data_move_warp_per_block=4,
                               sizeof_value=self.dtype.itemsize, is_complex_matrix=int(self.dtype.kind == "c"),
                               mpi_recv_buffer_level=0.5)
        self._h = C.c_void_p()
        self.L.pangulu_init(self.n, int(self._colptr[-1]), self._colptr.ctypes.data, self._rowidx.ctypes.data,
                            self._value.ctypes.data, C.byref(opt), C.byref(self._h))
        self.nb = int(self.info("nb"))

    # -- reference API -----------------------------------------------------------------------------
    def gstrf(self) -> float:
        opt = C.c_char(0)
        self.L.pangulu_gstrf(C.byref(opt), C.byref(self._h))
        return self.info("time_gstrf")

    def gstrs(self, rhs: np.ndarray) -> np.ndarray:
        x = np.array(rhs, dtype=self.dtype, copy=True)
        opt = C.c_char(0)
        self.L.pangulu_gstrs(x.ctypes.data, C.byref(opt), C.byref(self._h))
        return x

    def gssv(self, rhs: np.ndarray) -> np.ndarray:
        self.gstrf()
        return self.gstrs(rhs)

    def finalize(self) -> None:
        if self._h:
            self.L.pangulu_finalize(C.byref(self._h))
            self._h = C.c_void_p()

    # -- driver extras ------------------------------------------------------------------------------
    def info(self, key: str) -> float:
        return float(self.L.pangulu_sm100_driver_info(self._h, key.encode()))

    def reset_values(self, also_device: bool = True) -> None:
        self.L.pangulu_sm100_driver_reset_values(self._h, int(also_device))

    def update_values(self, values: np.ndarray) -> float:
        """New numerical values on the pattern the solver was built for (the data array of the sorted CSC matrix given to
        the constructor): nnz(A) values travel to the GPU and are scattered into the block value region by a kernel
        (preprocessing on the device, include/pangulu_sm100_driver.h).  Resident mode.  Returns seconds."""
        v = np.ascontiguousarray(values, dtype=self.dtype)
        assert v.shape == self._value.shape
        self._value = v
        return float(self.L.pangulu_sm100_driver_update_values(self._h, self._colptr.ctypes.data, self._rowidx.ctypes.data, v.ctypes.data))

    def device_mirrors(self, clear_first: bool = True) -> int:
        """Rebuild the CSR mirrors of the strictly-lower blocks on the GPU; returns the number of bytes in which the device's
        index region differs from the host-built one (0 = bit-identical)."""
        return int(self.L.pangulu_sm100_driver_device_mirrors(self._h, int(clear_first)))

    def gstrf_enqueue(self) -> None:
        """Resident mode: enqueue the factorisation on the module's stream without synchronising."""
        self.L.pangulu_sm100_driver_gstrf_enqueue(self._h)

    def synchronize(self) -> None:
        self.L.platform("synchronize")()

    def set_mode(self, mode: int) -> None:
        self.L.pangulu_sm100_driver_set_mode(self._h, int(mode))

    def download(self) -> None:
        self.L.pangulu_sm100_driver_download(self._h)

    # -- multi-GPU wiring ---------------------------------------------------------------------------------
    def mirror_table(self) -> np.ndarray:
        """int64[n, 7]: remote blocks this rank reads (brow, bcol, kind, value offset, flag offset, dense-copy offset, group)."""
        n = int(self.L.pangulu_sm100_driver_mirror_table(self._h, None, 0))
        out = np.zeros((max(n, 1), 7), dtype=np.int64)
        self.L.pangulu_sm100_driver_mirror_table(self._h, out.ctypes.data, n)
        return out[:n]

    def ipc_handles(self) -> bytes:
        buf = (C.c_ubyte * 192)()
        self.L.pangulu_sm100_driver_ipc_handles(self._h, buf)
        return bytes(buf)

    def connect(self, group=None) -> int:
        """Wire this rank's plan to all peers of a torch.distributed group (one rank per GPU): exchange CUDA IPC
        handles and mirror tables, register the NVLink pushes, install the barrier.  Returns the pushes registered."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.L.pangulu_sm100_driver_disconnect(self._h)   # (a reconnect after rebuild_plan: drop the old peer mappings)
        table = self.mirror_table()
        handles = self.ipc_handles() if not self._host_only else bytes(192)
        gathered = [None] * world
        dist.all_gather_object(gathered, (handles, table), group=group)
        made = 0
        for r, (hd, tb) in enumerate(gathered):
            if r == rank or len(tb) == 0:
                continue
            tb = np.ascontiguousarray(tb, dtype=np.int64)
            hb = (C.c_ubyte * 192).from_buffer_copy(hd)
            made += int(self.L.pangulu_sm100_driver_connect_peer(self._h, hb, tb.ctypes.data, len(tb)))
        self._barrier_cb = C.CFUNCTYPE(None)(lambda: dist.barrier(group))
        self.L.pangulu_sm100_driver_set_barrier(C.cast(self._barrier_cb, C.c_void_p))
        self.L.pangulu_sm100_driver_commit(self._h)   # (the handle keeps the barrier installed at this point)
        dist.barrier(group)
        return made

    def diag_groups(self, k: int) -> list[int]:
        """Index bounds of the independent groups of diagonal block k ([0, nb] when it is not split)."""
        b = (C.c_ushort * 5)()
        ng = int(self.L.pangulu_sm100_driver_diag_groups(self._h, int(k), b))
        return [int(b[i]) for i in range(ng + 1)]

    def save_factors(self, path: str) -> int:
        """Write the owned blocks (host copies: call download() first in resident mode) in PanguLU's packed block format."""
        n = int(self.L.pangulu_sm100_driver_save_factors(self._h, str(path).encode()))
        if n < 0:
            raise OSError(f"cannot write {path}")
        return n

    def load_factors(self, path: str) -> int:
        """Put the values of a factor file into this handle (same matrix, same nb): gstrs works without gstrf."""
        n = int(self.L.pangulu_sm100_driver_load_factors(self._h, str(path).encode()))
        if n < 0:
            raise OSError(f"cannot read {path}")
        return n

    def factor_check(self) -> float:
        return float(self.L.pangulu_sm100_driver_factor_check(self._h))

    def factor_check_distributed(self, group=None) -> float:
        """||L(U 1) - A 1|| / ||A 1|| with the factors spread over the ranks of a torch.distributed group, in the
        reference's own three passes (src/pangulu_numeric.c:1082-1341): every rank multiplies the blocks it owns
        (host copies: call download() first), the partial vectors are summed over the ranks.  Same value on every rank."""
        import torch
        import torch.distributed as dist

        N = int(self.info("block_length")) * self.nb
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"

        def allsum(v):
            t = torch.from_numpy(v).to(dev)
            dist.all_reduce(t, group=group)
            return t.cpu().numpy()

        ux = np.zeros(N)
        self.L.pangulu_sm100_driver_check_ux(self._h, ux.ctypes.data)
        ux = np.ascontiguousarray(allsum(ux))
        lux = np.zeros(N)
        self.L.pangulu_sm100_driver_check_lux(self._h, ux.ctypes.data, lux.ctypes.data)
        lux = allsum(lux)
        ax = np.zeros(N)
        n = int(self.L.pangulu_sm100_driver_check_ax(self._h, ax.ctypes.data))
        return float(np.linalg.norm(lux[:n] - ax[:n]) / np.linalg.norm(ax[:n]))

    def perms(self) -> dict:
        n = self.n
        g = lambda w, dt: _as_array(self.L.pangulu_sm100_driver_perm(self._h, w), n, dt).copy()
        return {"row_perm": g(0, np.uint32), "col_perm": g(1, np.uint32), "metis_perm": g(2, np.uint32),
                "row_scale": g(3, self.dtype), "col_scale": g(4, self.dtype)}

    def slot_ptr(self, i: int):
        return self.L.pangulu_sm100_driver_slot(self._h, i)

    def blocks(self, copy: bool = True, owned_only: bool = True) -> list[Block]:
        nb = self.nb
        out = []
        for i in range(int(self.info("num_slots"))):
            s = self.slot_ptr(i).contents
            owned = s.bin_id == 0
            if owned_only and not owned:
                continue
            cp = _as_array(s.columnpointer, nb + 1, np.uint32)
            nnz = int(cp[nb])
            arr = lambda p, cnt, dt: (_as_array(p, cnt, dt).copy() if copy else _as_array(p, cnt, dt))
            b = Block(s.brow_pos, s.bcol_pos, s.is_upper, owned, arr(s.columnpointer, nb + 1, np.uint32),
                      arr(s.rowindex, nnz, np.uint16), arr(s.value, nnz, self.dtype))
            if s.brow_pos > s.bcol_pos:
                b.rowptr = arr(s.rowpointer, nb + 1, np.uint32)
                b.colidx = arr(s.columnindex, nnz, np.uint16)
                b.csr2csc = arr(s.idx_of_csc_value_for_csr, nnz, np.uint32)
            out.append(b)
        return out

    def schedule(self):
        """(tasks as a numpy structured view, wave_ptr)."""
        wp = C.c_void_p()
        tp = self.L.pangulu_sm100_driver_tasks(self._h, C.byref(wp))
        nwave, ntask = int(self.info("nwave")), int(self.info("ntask"))
        wave_ptr = _as_array(wp.value, nwave + 1, np.uint64).copy()
        dt = np.dtype({"names": ["row", "col", "kernel_id", "task_level", "compare_flag", "opdst", "op1", "op2"],
                       "formats": [np.uint32, np.uint32, np.int16, np.uint32, np.int64, np.uint64, np.uint64, np.uint64],
                       "offsets": [0, 4, 8, 12, 16, 24, 32, 40], "itemsize": 48})
        tasks = _as_array(tp, ntask, dt) if ntask else np.zeros(0, dtype=dt)
        return tasks, wave_ptr

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass
