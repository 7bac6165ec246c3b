"""Run the unmodified reference (oracle/_ref) on one matrix in a FRESH process and save what it produced.

The reference keeps process-global state and cannot be initialised twice in one process, so every
factorisation gets its own interpreter:

    python -m oracle.ref_run in.npz out.npz --nb 64 [--vtype r64] [--perf] [--no-solve] [--no-blocks]

in.npz : indptr, indices, data, n (CSC).  out.npz: perms, blocks before/after gstrf (flattened), x, times,
flop.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def _flatten(blocks, prefix, out):
    out[prefix + "meta"] = np.array([[b.brow, b.bcol, b.is_upper, len(b.value)] for b in blocks], dtype=np.int64).reshape(-1, 4)
    out[prefix + "colptr"] = np.concatenate([b.colptr for b in blocks]) if blocks else np.zeros(0, np.uint32)
    out[prefix + "rowidx"] = np.concatenate([b.rowidx for b in blocks]) if blocks else np.zeros(0, np.uint16)
    out[prefix + "value"] = np.concatenate([b.value for b in blocks]) if blocks else np.zeros(0)


def unflatten(z, prefix, nb):
    """Inverse of _flatten: list of simple block objects (CSR mirrors rebuilt on demand by the caller)."""
    from types import SimpleNamespace

    meta = z[prefix + "meta"]
    colptr, rowidx, value = z[prefix + "colptr"], z[prefix + "rowidx"], z[prefix + "value"]
    out, o = [], 0
    for i, (brow, bcol, is_upper, nnz) in enumerate(meta):
        b = SimpleNamespace(brow=int(brow), bcol=int(bcol), is_upper=int(is_upper),
                            colptr=np.ascontiguousarray(colptr[i * (nb + 1):(i + 1) * (nb + 1)]),
                            rowidx=np.ascontiguousarray(rowidx[o:o + nnz]), value=np.ascontiguousarray(value[o:o + nnz]),
                            rowptr=None, colidx=None, csr2csc=None)
        o += int(nnz)
        out.append(b)
    return out


def add_csr_mirror(b, nb):
    """rowptr/colidx/csr2csc of a CSC block (same result as pangulu_transpose_struct_with_valueidx_inblock)."""
    nnz = int(b.colptr[nb])
    cols = np.repeat(np.arange(nb, dtype=np.uint16), np.diff(b.colptr.astype(np.int64)))
    order = np.lexsort((cols, b.rowidx[:nnz]))
    b.colidx = np.ascontiguousarray(cols[order])
    b.csr2csc = np.ascontiguousarray(order.astype(np.uint32))
    cnt = np.bincount(b.rowidx[:nnz], minlength=nb)
    b.rowptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    return b


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("inp")
    ap.add_argument("out")
    ap.add_argument("--nb", type=int, required=True)
    ap.add_argument("--vtype", default="r64")
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-blocks", action="store_true")
    ap.add_argument("--nthread", type=int, default=1)
    ap.add_argument("--flavour", default="cpu", help="cpu: reference CPU platform; gpu: reference host code on top of the sm_100a module; "
                    "plan: the same plus the integration overlay (pangulu_gstrf -> this module's plan path)")
    a = ap.parse_args(argv)
    sys.path.insert(0, _ROOT)
    from oracle.ref import RefLib

    z = np.load(a.inp)
    n = int(z["n"])
    R = RefLib(a.vtype, perf=a.perf, flavour=a.flavour)
    f = R.factor(n, z["indptr"], z["indices"], z["data"], a.nb, nthread=a.nthread)
    out = {"n": n, "nb": f.nb, "symbolic_nnz": f.symbolic_nnz, "has_metis": int(R.has_metis), "has_mc64": int(R.has_mc64)}
    out.update(f.perms())
    if not a.no_blocks:
        _flatten(f.blocks(), "pre_", out)
    t = f.gstrf()
    out["t_gstrf"] = t
    if a.flavour.startswith("plan"):
        # the overlay's pangulu_gstrf ran this module's plan path: its phase times, and a second (warm: schedule and plan cached)
        # factorisation after the values of A were put back into PanguLU's blocks
        import ctypes as C

        L = R.lib
        L.pangulu_sm100_reference_gstrf_time.restype = C.c_double
        L.pangulu_sm100_reference_gstrf_time.argtypes = [C.c_void_p, C.c_int]
        slots = L.refh_slot(f.ctx, 0)
        out["plan_times"] = np.array([L.pangulu_sm100_reference_gstrf_time(slots, k) for k in range(4)])
    out["flop"] = f.flop()
    out["blas"] = R.blas_backend
    if not a.no_blocks:
        _flatten(f.blocks(), "post_", out)
    if not a.no_solve:
        import scipy.sparse as sp

        A = sp.csc_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
        b = A @ np.ones(n, dtype=A.dtype)
        x = f.gstrs(b)
        out["x"] = x
        out["resid"] = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    np.savez(a.out, **out)


def run_reference(A, nb, vtype="r64", perf=False, solve=True, blocks=True, nthread=1, quiet=True, flavour="cpu", ranks=1):
    """Factor scipy CSC matrix A with the reference in a subprocess; returns the loaded npz dict (of rank 0).

    ranks > 1: that many reference processes ("MPI ranks" of the shim oracle/shims/mpi_local.c: messages travel through
    POSIX shared memory) factor the matrix together, as `mpirun -np ranks` would; `t_gstrf` is then the maximum over the
    ranks.  Only the timing is meaningful in that mode (every rank holds its own blocks)."""
    import scipy.sparse as sp

    A = sp.csc_matrix(A)
    A.sort_indices()
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.npz")
        np.savez(inp, indptr=A.indptr.astype(np.uint64), indices=A.indices.astype(np.uint32), data=A.data, n=A.shape[0])
        # The reference leaves the value arrays of some fill-in entries uninitialised (malloc without memset in its
        # block construction); on small problems recycled heap memory then shows up as garbage, even NaN, in the
        # factors.  glibc's MALLOC_PERTURB_=255 makes malloc hand out zero-filled memory (byte ^ 0xff), which is what
        # the algorithm assumes -- the reference sources stay untouched.
        env = dict(os.environ, MALLOC_PERTURB_="255")
        if ranks > 1:
            from oracle.ref import ref_lib_path

            session = "pgl_%d_%d" % (os.getpid(), int(time.time() * 1e3) % 1000000)
            mk = ("import ctypes,sys; sys.exit(ctypes.CDLL(%r).pangulu_shim_create_session(%r, %d))"
                  % (ref_lib_path(vtype, perf, flavour), session.encode(), ranks))
            if subprocess.run([sys.executable, "-c", mk], cwd=_ROOT).returncode != 0:
                raise RuntimeError("could not create the shared-memory session of the MPI shim")
            env.update(PANGULU_SHIM_NP=str(ranks), PANGULU_SHIM_SESSION=session, OMP_NUM_THREADS="1")
        procs, outs = [], []
        for r in range(ranks):
            out = os.path.join(td, f"out{r}.npz")
            outs.append(out)
            cmd = [sys.executable, "-m", "oracle.ref_run", inp, out, "--nb", str(nb), "--vtype", vtype, "--nthread", str(nthread),
                   "--flavour", flavour]
            if perf:
                cmd.append("--perf")
            if not solve or ranks > 1:
                cmd.append("--no-solve")
            if not blocks or ranks > 1:
                cmd.append("--no-blocks")
            e = dict(env, PANGULU_SHIM_RANK=str(r)) if ranks > 1 else env
            procs.append(subprocess.Popen(cmd, cwd=_ROOT, env=e, stdout=subprocess.PIPE if quiet else None, stderr=subprocess.STDOUT if quiet else None))
        for p in procs:
            so = p.communicate()[0]
            if p.returncode != 0:
                for q in procs:
                    if q.poll() is None:
                        q.kill()
                raise RuntimeError("reference run failed:\n" + (so.decode(errors="replace") if so else ""))
        res = None
        tmax = 0.0
        for out in outs:
            with np.load(out, allow_pickle=False) as z:
                d = {k: z[k] for k in z.files}
            tmax = max(tmax, float(d["t_gstrf"]))
            if res is None:
                res = d
        res["t_gstrf"] = np.float64(tmax)
        return res


if __name__ == "__main__":
    main()
