"""Development probe: device time of the GETRF variants on dense diagonal blocks (nb x nb, all entries stored), through
the platform C ABI (hybrid_batched with one GETRF task per block), and agreement with the oracle."""
import ctypes as C, sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from types import SimpleNamespace
from helpers import DeviceBlocks, rel_err
from pangulu_b200 import _lib
from oracle.oracle import Oracle

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = _lib.load("r64")
rng = np.random.default_rng(1)
M = rng.uniform(-1, 1, (nb, nb)) + np.eye(nb) * nb
# lower: strict lower CSC; upper: CSR incl. diagonal (diagonal first = ascending columns starting at the diagonal)
lcp = np.zeros(nb + 1, np.uint32); lri = []; lv = []
for c in range(nb):
    rows = np.arange(c + 1, nb); lri.append(rows); lv.append(M[rows, c]); lcp[c + 1] = lcp[c] + len(rows)
ucp = np.zeros(nb + 1, np.uint32); uci = []; uv = []
for r in range(nb):
    cols = np.arange(r, nb); uci.append(cols); uv.append(M[r, cols]); ucp[r + 1] = ucp[r] + len(cols)
mk = lambda brow, up, cp, idx, val: SimpleNamespace(brow=brow, bcol=brow, is_upper=up, colptr=cp.copy(), rowidx=np.concatenate(idx).astype(np.uint16), value=np.concatenate(val).astype(np.float64), rowptr=None, colidx=None, csr2csc=None)
for mode, name in ((3, "dense-promoted"), (1, "global-memory sparse")):
    L.pangulu_sm100_set_option(b"getrf_mode", mode)
    D = DeviceBlocks("r64", nb)
    slots, tasks = [], []
    for k in range(nblk):
        lo, up = mk(k, 0, lcp, lri, lv), mk(k, 1, ucp, uci, uv)
        s_lo, s_up = D.slot(lo), D.slot(up)
        D.link(s_lo, s_up)
        slots.append((s_lo, s_up, lo, up))
        tasks.append(_lib.Task(row=k, col=k, kernel_id=_lib.TASK_GETRF, task_level=k, opdst=C.addressof(s_lo)))
    arr = (_lib.Task * len(tasks))(*tasks)
    L.pangulu_sm100_set_option(b"eager_download", 0)
    D.P("hybrid_batched")(nb, len(tasks), C.addressof(arr)); D.P("synchronize")()
    got_l = D.fetch(slots[0][0], len(slots[0][2].value)); got_u = D.fetch(slots[0][1], len(slots[0][3].value))
    wl, wu = copy.deepcopy(slots[0][2]), copy.deepcopy(slots[0][3])
    Oracle("r64").getrf(nb, wl, wu)
    err = max(rel_err(got_l, wl.value), rel_err(got_u, wu.value))
    # timing: one block at a time (latency of ONE factorisation on one SM), values reset before every call
    one = (_lib.Task * 1)(tasks[0])
    ts = []
    for rep in range(5):
        D.P("memcpy")(slots[0][0].d_value, slots[0][2].value.ctypes.data, slots[0][2].value.nbytes, 0)
        D.P("memcpy")(slots[0][1].d_value, slots[0][3].value.ctypes.data, slots[0][3].value.nbytes, 0)
        L.pangulu_sm100_stats_reset(); L.pangulu_sm100_set_option(b"profile", 1)
        D.P("hybrid_batched")(nb, 1, C.addressof(one)); D.P("synchronize")()
        ts.append(L.pangulu_sm100_stats_ms(0)); L.pangulu_sm100_set_option(b"profile", 0)
    print(f"nb {nb} GETRF {name:22s}: one block on one SM {min(ts)*1e3:8.1f} us   max rel. error vs oracle {err:.2e}")
    D.close()
L.pangulu_sm100_set_option(b"getrf_mode", 0); L.pangulu_sm100_set_option(b"eager_download", 1)
