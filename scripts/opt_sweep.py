"""Development probe: one init, several platform option settings, a few steps each (resident plan), 100^3 by default."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pangulu_b200 import PanguLU, matrices, _lib
size = int(os.environ.get("SIZE", "100")); nb = int(os.environ.get("NB", "256"))
settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in sys.argv[1:]] or [{}]
L = _lib.load("r64")
s = PanguLU(matrices.poisson3d(size), nb=nb, mode=1)
flop = s.info("flop")
for st in settings:
    for k, v in st.items():
        assert L.pangulu_sm100_set_option(k.encode(), int(v)) == 0, k
    t0 = time.time(); L.pangulu_sm100_driver_rebuild_plan(s._h); tp = time.time() - t0
    ts = []
    for it in range(3):
        s.reset_values(True)
        ts.append(s.gstrf())
    s.download()
    print(json.dumps({"opts": st, "plan_s": round(tp, 1), "gstrf_ms": [round(t * 1e3, 1) for t in ts], "gflops": round(flop / min(ts) / 1e9, 1),
                      "check": s.factor_check(), "dense_GB": round(s.info("plan_dense_bytes") / 1e9, 2), "ct_tasks": s.info("plan_ct_tasks")}), flush=True)
