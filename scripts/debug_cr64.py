"""Development probe: locate a wrong factor at production block size (nb 256) for any value type by toggling the plan options.
Prints the factor check, the count of non-finite blocks and the first block that is off, per option set."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pangulu_b200 import PanguLU, matrices, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--vtype", default="cr64"); ap.add_argument("--n", type=int, default=3000); ap.add_argument("--nb", type=int, default=256)
ap.add_argument("--gen", default="circuit"); ap.add_argument("--oracle", type=int, default=0)
a = ap.parse_args()
dtype = _lib.VALUE_DTYPES[a.vtype]
A = matrices.circuit_like(a.n, dtype=dtype) if a.gen == "circuit" else getattr(matrices, a.gen)(a.n).astype(dtype)
L = _lib.load(a.vtype)
DEFAULTS = {"executor": 1, "compact": 1, "cpanel": 1, "getrf_mode": 0, "compact_threshold": 0, "compact_block_threshold": 0}
SETS = [{}, {"executor": 0}, {"compact": 0}, {"compact": 0, "executor": 0}, {"getrf_mode": 2}, {"getrf_mode": 2, "compact": 0}]
want = None
if a.oracle:
    from types import SimpleNamespace
    from oracle.oracle import Oracle
    s = PanguLU(A, nb=a.nb, mode=0)
    want = [SimpleNamespace(**vars(b)) for b in s.blocks(copy=True)]
    Oracle(a.vtype).block_lu(a.nb, want)
    s.finalize()
for opts in SETS:
    for k, v in {**DEFAULTS, **opts}.items():
        assert L.pangulu_sm100_set_option(k.encode(), v) == 0, k
    s = PanguLU(A, nb=a.nb, mode=1)
    try:
        s.gstrf(); s.download()
        blocks = s.blocks(copy=True)
        bad = [(b.brow, b.bcol, int(b.is_upper)) for b in blocks if b.value.size and not np.isfinite(b.value).all()]
        line = f"{opts!s:40s} factor_check {s.factor_check():.3e}  non-finite blocks {len(bad)} of {len(blocks)} first {bad[:4]}"
        if want is not None:
            errs = [(float(np.abs(g.value - w.value).max() / max(np.abs(w.value).max(), 1e-300)) if w.value.size else 0.0, g.brow, g.bcol, int(g.is_upper))
                    for g, w in zip(blocks, want)]
            off = [e for e in errs if not (e[0] <= 1e-10)]
            line += f" | vs oracle: worst {max(e[0] for e in errs):.2e}, blocks off {len(off)} first {sorted(off, key=lambda e: (min(e[1], e[2]), e[1], e[2]))[:4]}"
        print(line, flush=True)
        print("   stats: " + " ".join(f"{k}={int(s.info(k))}" for k in ("nwave", "ntask", "n_ssssm")), flush=True)
    finally:
        s.finalize()
