"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref, built in place
from /root/reference by oracle/Makefile).  Run where the reference tree exists:

    python tests/golden/make_golden.py

Outputs (all small):
  abi_offsets.json            field offsets of pangulu_storage_slot_t / pangulu_task_t in the reference's GPU build
  <case>.npz                  for each case: permutation, blocks before and after pangulu_gstrf (pattern+values),
                              solution of A x = A 1, flop count of the reference's own model (-DPANGULU_PERF build)
  Trefethen_20b_pattern.npz   the reference's shipped fixture examples/Trefethen_20b.mtx as CSC arrays
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_run import run_reference  # noqa: E402
from pangulu_b200 import matrices  # noqa: E402

CASES = {
    "trefethen_nb4": (lambda: matrices.trefethen("/root/reference/examples/Trefethen_20b.mtx"), 4),
    "trefethen_nb10": (lambda: matrices.trefethen("/root/reference/examples/Trefethen_20b.mtx"), 10),
    "trefethen_nb19": (lambda: matrices.trefethen("/root/reference/examples/Trefethen_20b.mtx"), 19),
    "poisson3d_6_nb16": (lambda: matrices.poisson3d(6), 16),
    "poisson3d_6_nb24": (lambda: matrices.poisson3d(6), 24),
    "laplace2d_12_nb8": (lambda: matrices.laplace2d(12), 8),
    "stencil27_5_nb32": (lambda: matrices.stencil27(5), 32),
}


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpangulu_ref_abi_gpu.so"))
    off = (C.c_longlong * 30)()
    lib.refabi_offsets(off)
    names = ["slot.brow_pos", "slot.bcol_pos", "slot.columnpointer", "slot.rowindex", "slot.value", "slot.rowpointer", "slot.columnindex",
             "slot.idx_of_csc_value_for_csr", "slot.data_status", "slot.related_block", "slot.is_upper", "slot.bin_id", "slot.slot_idx",
             "slot.task_queue", "slot.d_columnpointer", "slot.d_rowindex", "slot.d_value", "slot.d_rowpointer", "slot.d_columnindex",
             "slot.d_idx_of_csc_value_for_csr", "sizeof(slot)", "task.row", "task.col", "task.kernel_id", "task.task_level",
             "task.compare_flag", "task.opdst", "task.op1", "task.op2", "sizeof(task)"]
    with open(os.path.join(HERE, "abi_offsets.json"), "w") as f:
        json.dump({n: int(v) for n, v in zip(names, off)}, f, indent=1)
    A = matrices.trefethen("/root/reference/examples/Trefethen_20b.mtx")
    np.savez_compressed(os.path.join(HERE, "Trefethen_20b_pattern.npz"), indptr=A.indptr, indices=A.indices, data=A.data, n=A.shape[0])
    for name, (make, nb) in CASES.items():
        A = make()
        z = run_reference(A, nb, perf=True)
        # The reference does not zero the value arrays of every fill-in entry (uninitialised heap memory shows
        # up as garbage in small problems -- see DESIGN.md "reference defects").  A golden file is only kept if
        # the reference's blocks BEFORE gstrf equal "values of A scattered into a zero-filled pattern".
        from pangulu_b200 import PanguLU
        mine = PanguLU(A, nb=nb, host_only=True)
        clean = np.array_equal(np.concatenate([b.value for b in mine.blocks()]), z["pre_value"])
        mine.finalize()
        if not clean:
            print(name, "SKIPPED: reference started from uninitialised fill-in values in this process")
            continue
        keep = {k: z[k] for k in ("n", "nb", "symbolic_nnz", "metis_perm", "row_perm", "col_perm", "pre_meta", "pre_colptr", "pre_rowidx",
                                  "pre_value", "post_meta", "post_colptr", "post_rowidx", "post_value", "x", "resid", "flop", "has_metis", "has_mc64")}
        keep["A_indptr"], keep["A_indices"], keep["A_data"] = A.indptr, A.indices, A.data
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **keep)
        print(name, "flop", int(z["flop"]), "resid", float(z["resid"]))


if __name__ == "__main__":
    main()
