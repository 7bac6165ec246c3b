"""Development probe: time gstrf (resident plan) on one synthetic matrix and print per-kernel-class times."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pangulu_b200 import PanguLU, matrices

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="p3"); ap.add_argument("--size", type=int, default=32)
ap.add_argument("--nb", type=int, default=128); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--mode", type=int, default=1); ap.add_argument("--profile", type=int, default=1)
ap.add_argument("--opt", action="append", default=[]); ap.add_argument("--schedule", type=int, default=2); ap.add_argument("--lookahead", type=int, default=2)
a = ap.parse_args()
A = {"p3": matrices.poisson3d, "p2": matrices.laplace2d, "s27": matrices.stencil27}[a.case](a.size)
from pangulu_b200 import _lib
L = _lib.load("r64")
for kv in a.opt:  # platform options must be in place before the plan is built
    k, v = kv.split("="); assert L.pangulu_sm100_set_option(k.encode(), int(v)) == 0, k
t = time.time(); s = PanguLU(A, nb=a.nb, mode=a.mode, schedule=a.schedule, lookahead=a.lookahead); t_init = time.time() - t
print(f"n={A.shape[0]} nnz={A.nnz} nb={a.nb} init {t_init:.2f}s  " + " ".join(f"{k}={s.info('time_'+k):.2f}" for k in ("reorder","symbolic","blocks","schedule","work_model")))
print(f"symbolic_nnz={s.info('symbolic_nnz'):.3e} nwave={int(s.info('nwave'))} ntask={int(s.info('ntask'))} ssssm={int(s.info('n_ssssm'))} flop={s.info('flop'):.4e} value_MB={s.info('value_bytes')/1e6:.1f}")
for it in range(a.steps):
    if it: s.reset_values(True)
    L.pangulu_sm100_stats_reset()
    L.pangulu_sm100_set_option(b"profile", 0)
    t = s.gstrf()
    print(f"step {it}: gstrf {t*1e3:.2f} ms  {s.info('flop')/t/1e9:.1f} GFLOP/s  launches " + " ".join(str(L.pangulu_sm100_stats_launches(c)) for c in range(7)))
if a.profile:
    s.reset_values(True); L.pangulu_sm100_stats_reset(); L.pangulu_sm100_set_option(b"profile", 1)
    t = s.gstrf()
    names = ["getrf", "panel", "-", "ssssm_sparse", "ssssm_tile", "aux", "mega"]
    print("profiled (serialised) ms per class: " + " ".join(f"{names[c]}={L.pangulu_sm100_stats_ms(c):.2f}" for c in range(7)) + f" total {t*1e3:.2f}")
    for c, nm in ((0, "getrf"), (1, "tstrf"), (1, "gessm"), (3, "ssssm")):
        pass
    print("flops: " + " ".join(f"{k}={s.info('flop_'+k):.3e}" for k in ("getrf","tstrf","gessm","ssssm")))
    print("bytes: " + " ".join(f"{k}={s.info('bytes_'+k):.3e}" for k in ("getrf","tstrf","gessm","ssssm")))
    L.pangulu_sm100_set_option(b"profile", 0)
if a.mode == 1: s.download()
print("factor check", s.factor_check())
b = A @ np.ones(A.shape[0]); x = s.gstrs(b); print("resid", np.linalg.norm(A @ x - b) / np.linalg.norm(b))
