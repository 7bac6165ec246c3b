"""Workload statistics of the compact-tile SSSSM path from the host-only structure pass (no GPU needed)."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from pangulu_b200 import PanguLU, matrices, _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kind = sys.argv[3] if len(sys.argv) > 3 else "poisson3d"
A = {"poisson3d": matrices.poisson3d, "laplace2d": matrices.laplace2d, "stencil27": matrices.stencil27}[kind](size)
t0 = time.time()
s = PanguLU(A, nb=nb, host_only=True)
print("init %.1fs" % (time.time() - t0), "flop %.3e" % s.info("flop"), "blocks", int(s.info("num_slots")), "nwave", int(s.info("nwave")))
nslot = int(s.info("num_slots"))
base = C.addressof(s.slot_ptr(0).contents)
tasks, wave_ptr = s.schedule()
# per-slot pattern summaries
rows_cnt = np.zeros(nslot, np.int32); cols_cnt = np.zeros(nslot, np.int32); nnz = np.zeros(nslot, np.int64)
kmask_r = np.zeros(nslot, np.uint32); kmask_c = np.zeros(nslot, np.uint32)
colmask = np.zeros((nslot, nb), bool)
for i in range(nslot):
    sl = s.slot_ptr(i).contents
    cp = np.ctypeslib.as_array(C.cast(sl.columnpointer, C.POINTER(C.c_uint32)), (nb + 1,))
    n = int(cp[nb]); nnz[i] = n
    if n == 0: continue
    ri = np.ctypeslib.as_array(C.cast(sl.rowindex, C.POINTER(C.c_uint16)), (n,))
    present = np.zeros(nb, bool); present[ri] = True
    rows_cnt[i] = present.sum()
    cl = np.diff(np.concatenate([[0], cp[1:]]).astype(np.int64)) > 0
    cl[0] = cp[1] > 0
    colmask[i] = cl
    cols_cnt[i] = cl.sum()
    km = 0
    for ch in range(nb // 16):
        if present[ch*16:(ch+1)*16].any(): km |= 1 << ch
    kmask_r[i] = km   # as B operand: rows = k
    km = 0
    for ch in range(nb // 16):
        if cl[ch*16:(ch+1)*16].any(): km |= 1 << ch
    kmask_c[i] = km   # as A operand: cols = k
ss = tasks[tasks["kernel_id"] == 4]
a = ((ss["op1"] - base) // 144).astype(np.int64); b = ((ss["op2"] - base) // 144).astype(np.int64); d = ((ss["opdst"] - base) // 144).astype(np.int64)
m = rows_cnt[a]; n = cols_cnt[b]
pc = np.array([bin(int(x)).count("1") for x in (kmask_c[a] & kmask_r[b])])
dm = ((m + 7) // 8) * ((n + 7) // 8) * pc * 4      # DMMA instr if whole block columns at once (ignores 32-col tiling)
print("ssssm tasks", len(ss), "dests", len(np.unique(d)))
print("DMMA total(untiled) %.3e  -> flop %.3e" % (dm.sum(), dm.sum() * 512))
w = dm / dm.sum()
def wq(x, w, qs=(0.1, 0.25, 0.5, 0.75, 0.9)):
    o = np.argsort(x); cw = np.cumsum(w[o]); return [int(x[o][np.searchsorted(cw, q)]) for q in qs]
print("count-weighted  m' quantiles", np.percentile(m, [10,25,50,75,90]).astype(int), " n'", np.percentile(n, [10,25,50,75,90]).astype(int), " chunks", np.percentile(pc, [10,25,50,75,90]).astype(int))
print("DMMA-weighted   m' quantiles", wq(m, w), " n'", wq(n, w), " chunks", wq(pc, w))
# pairs per destination
u, cnt = np.unique(d, return_counts=True)
print("pairs per dest: mean %.1f median %d p90 %d max %d" % (cnt.mean(), np.median(cnt), np.percentile(cnt, 90), cnt.max()))
# DMMA-weighted pairs per dest
dsum = np.zeros(nslot); np.add.at(dsum, d, dm)
print("DMMA share of dests with >=16 pairs: %.3f, >=64: %.3f" % (dsum[u[cnt >= 16]].sum() / dm.sum(), dsum[u[cnt >= 64]].sum() / dm.sum()))
# operand bytes if A streamed once per (pair, 32-col tile) vs once per pair
ld = ((m + 7) // 8) * 8 + 4
tiles32 = np.maximum(1, (cols_cnt[d] + 31) // 32)
print("A bytes per step, per pair once: %.3e ; x dest 32-col tiles: %.3e ; B bytes: %.3e" % ((pc * 16 * ld * 8).sum(), (pc * 16 * ld * 8 * tiles32).sum(), (pc * n * 20 * 8).sum()))
for thr in (8, 16, 32, 64, 128, 256):
    sel = n <= thr
    print("  pairs with n' <= %3d: count share %.3f DMMA share %.3f" % (thr, sel.mean(), w[sel].sum()))
for thr in (8, 16, 32, 64, 128, 256):
    sel = m <= thr
    print("  pairs with m' <= %3d: count share %.3f DMMA share %.3f" % (thr, sel.mean(), w[sel].sum()))
s.finalize()
