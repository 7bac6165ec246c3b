"""Development probe: per-task timeline of the persistent executor (option trace=1)."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pangulu_b200 import PanguLU, matrices, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=48); ap.add_argument("--nb", type=int, default=256)
ap.add_argument("--schedule", type=int, default=2); ap.add_argument("--lookahead", type=int, default=2)
ap.add_argument("--opt", action="append", default=[]); ap.add_argument("--out", default="")
a = ap.parse_args()
L = _lib.load("r64")
L.pangulu_sm100_set_option(b"trace", 1); L.pangulu_sm100_set_option(b"executor", 1)
for kv in a.opt:
    k, v = kv.split("="); assert L.pangulu_sm100_set_option(k.encode(), int(v)) == 0
s = PanguLU(matrices.poisson3d(a.size), nb=a.nb, mode=1, schedule=a.schedule, lookahead=a.lookahead)
s.gstrf(); s.reset_values(True)
if os.environ.get('CT_PROF'):
    s.info('plan_ct_prof_on' + os.environ['CT_PROF'])
    if os.environ.get('CT_PROF_MAX'): s.info('plan_ct_prof_max' + os.environ['CT_PROF_MAX'])
t = s.gstrf()
n = L.pangulu_sm100_driver_trace(s._h, None, 0)
buf = np.zeros((n, 4), dtype=np.uint64)
L.pangulu_sm100_driver_trace(s._h, buf.ctypes.data, n)
if a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    np.save(a.out, buf)
kind = (buf[:, 0] & 0xff).astype(int); npairs = (buf[:, 0] >> 8).astype(int); t0 = buf[:, 1].min()
fetch, start, end = [(buf[:, i] - t0).astype(np.int64) / 1e3 for i in (1, 2, 3)]
names = ["tile", "sparse", "getrf_smem", "getrf_dense", "getrf_global", "panel", "ctile", "cpanel", "ddens"]
print(f"gstrf {t*1e3:.1f} ms, {n} tasks, span {end.max()/1e3:.1f} ms")
for k in range(9):
    m = kind == k
    if m.any():
        run, wait = end[m] - start[m], start[m] - fetch[m]
        print(f"{names[k]:12s} n={m.sum():7d} run total {run.sum()/1e3:9.1f} ms mean {run.mean():8.1f} us max {run.max():9.1f} | wait total {wait.sum()/1e3:9.1f} ms mean {wait.mean():8.1f} us")
st = {k: s.info("plan_" + k) for k in ("ct_tasks", "ct_pairtiles", "ct_chunks", "ct_dmma", "ct_bytes", "ct_msum", "ct_ntsum", "dense_bytes", "dense_blocks")}
if st["ct_pairtiles"] > 0:
    print("compact tiles: tasks %.0f pair-tiles %.0f chunks/pair-tile %.2f mean m' %.1f mean nt %.1f | DMMA %.3e -> %.1f CTA-s at 16.1 cyc | operand bytes %.3e | dense copies %.0f blocks %.2f GB"
          % (st["ct_tasks"], st["ct_pairtiles"], st["ct_chunks"] / st["ct_pairtiles"], st["ct_msum"] / st["ct_pairtiles"], st["ct_ntsum"] / st["ct_pairtiles"],
             st["ct_dmma"], st["ct_dmma"] * 16.1 / 4 / 1.965e9, st["ct_bytes"], st["dense_blocks"], st["dense_bytes"] / 1e9))
if os.environ.get('CT_PROF'):
    names7 = ["prologue", "staging", "wait_operands", "compute", "flush", "begin_run", "epilogue", "other", "producer_wait_slot", "producer_issue"]
    for who, base in (("thread 0", 0), ("producer thread", 10)):
        v = [s.info("plan_ct_prof%d" % (base + k)) for k in range(10)]
        tot = sum(v)
        if tot > 0:
            print(f"compact tile phases, tasks with >= {os.environ['CT_PROF']} pairs ({who}, cycles): " + "  ".join(f"{n} {x/tot*100:.1f}%" for n, x in zip(names7, v)) + f"  total {tot/1.965e9:.1f} CTA-s")
    v = [s.info("plan_ct_prof%d" % (20 + k)) for k in range(7)]
    if sum(v) > 0:
        cn = ["clear+staging", "gather", "wait_factor", "tensor_update", "substitution", "scatter", "dense_copy"]
        print("compact panel phases (thread 0, cycles): " + "  ".join(f"{n} {x/sum(v)*100:.1f}%" for n, x in zip(cn, v)) + f"  total {sum(v)/1.965e9:.1f} CTA-s")
print(f"flop {s.info('flop'):.4e}  -> {s.info('flop')/t/1e9:.1f} GFLOP/s")
for k in (0, 6):
    m = kind == k
    if m.any():
        print(names[k], "tasks: pairs total", npairs[m].sum(), "mean", npairs[m].mean().round(2), "max", npairs[m].max(), "| us per pair-tile", ((end[m]-start[m]).sum()/npairs[m].sum()).round(2))
# efficiency of the compact tile tasks by task weight: DMMA instructions of the task at 16.1 cycles per instruction and SM
# sub-partition (4 per SM) against the time the task held its SM
wk = np.zeros(n, dtype=np.float32)
if L.pangulu_sm100_driver_task_work(s._h, wk.ctypes.data, n) == n:
    m = kind == 6
    d, rt = wk[m].astype(np.float64), (end[m] - start[m]) * 1e-6
    floor = d * 16.1 / 4 / 1.965e9
    edges = [0, 1e3, 1e4, 1e5, 3e5, 1e6, 3e6, 1e7, 1e12]
    print("compact tile tasks by DMMA count: [lo, hi)  tasks  share of DMMA  busy CTA-s  DMMA floor CTA-s  efficiency  mean pairs")
    for lo, hi in zip(edges[:-1], edges[1:]):
        b = (d >= lo) & (d < hi)
        if b.any():
            print(f"  [{lo:8.0e},{hi:8.0e}) {b.sum():8d} {d[b].sum()/d.sum():6.3f} {rt[b].sum():9.2f} {floor[b].sum():9.2f} {floor[b].sum()/rt[b].sum():6.3f} {npairs[m][b].mean():8.1f}")
    print(f"  all: busy {rt.sum():.2f} CTA-s, floor {floor.sum():.2f} CTA-s, efficiency {floor.sum()/rt.sum():.3f}")
busy = (end - start).sum() / 1e3; waiting = (start - fetch).sum() / 1e3
print(f"sum over CTAs: busy {busy:.1f} ms, waiting {waiting:.1f} ms, capacity 148 x span = {148*end.max()/1e3:.1f} ms")
# critical chain through getrf tasks
g = np.where((kind >= 2) & (kind <= 4))[0]
gs, ge = start[g], end[g]
o = np.argsort(ge)
print("getrf: last 10 ends (ms):", np.round(ge[o][-10:] / 1e3, 2))
print("getrf completion gaps (us) percentiles 50/90/99:", np.percentile(np.diff(np.sort(ge)), [50, 90, 99]).round(1))

# ---- the tail: what runs between the last GETRF tasks (the dense front of the top separator), and how busy the machine is
last = o[-14:]
print("tail of the factorisation: per interval between consecutive GETRF completions -> SM utilisation and task mix")
for a_, b_ in zip(last[:-1], last[1:]):
    t0_, t1_ = ge[a_], ge[b_]
    inside = (end > t0_) & (start < t1_)
    ov = np.minimum(end[inside], t1_) - np.maximum(start[inside], t0_)
    util = ov.sum() / (148 * max(t1_ - t0_, 1e-9))
    mix = {names[k]: (int(((kind == k) & inside).sum()), round(float(ov[kind[inside] == k].sum()) / 1e3, 2)) for k in range(9) if ((kind == k) & inside).any()}
    gi = g[b_]
    print(f"  [{t0_/1e3:8.2f}, {t1_/1e3:8.2f}] ms ({(t1_-t0_)/1e3:5.2f} ms) util {util:5.3f}  getrf run {(end[gi]-start[gi])/1e3:5.2f} ms wait {(start[gi]-fetch[gi])/1e3:5.2f} ms  tasks(count, busy ms): {mix}")
