"""On-box comparator: the reference's OWN CUDA platform (0201000, compiled unmodified for sm_100 by oracle/Makefile) against its CPU
platform and against this module, same matrix, same block size.  Prints one JSON line per size (committed under profiles/)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.ref import ref_available
from oracle.ref_run import run_reference, unflatten
from pangulu_b200 import PanguLU, matrices

nb = 256
for size in [int(x) for x in (sys.argv[1:] or ["32", "48"])]:
    A = matrices.poisson3d(size)
    cpu = run_reference(A, nb, flavour="cpu")
    out = {"workload": f"poisson3d_7pt_{size}^3_r64_metis_nb{nb}", "n": int(A.shape[0]), "reference_cpu_1rank_gstrf_s": float(cpu["t_gstrf"])}
    if ref_available("r64", flavour="cuda"):
        try:
            cu = run_reference(A, nb, flavour="cuda")
            want, got = unflatten(cpu, "post_", nb), unflatten(cu, "post_", nb)
            worst = max(float(np.abs(g.value - w.value).max() / max(np.abs(w.value).max(), 1e-300)) if w.value.size else 0.0 for g, w in zip(got, want))
            out.update({"reference_cuda_platform_gstrf_s": float(cu["t_gstrf"]), "reference_cuda_platform_max_block_rel_err_vs_cpu": worst,
                        "reference_cuda_platform_resid": float(cu["resid"])})
        except Exception as e:  # the stock platform may fail on this GPU generation: say so, do not hide it
            out["reference_cuda_platform_error"] = str(e)[-300:]
    s = PanguLU(A, nb=nb, mode=2)
    s.gstrf(); t2 = s.gstrf() if False else s.info("time_gstrf")
    s.reset_values(True); t2 = s.gstrf()
    out["this_module_gstrf_host_buffers_s"] = float(t2)
    s.finalize()
    s = PanguLU(A, nb=nb, mode=1)
    s.gstrf(); s.reset_values(True)
    out["this_module_gstrf_resident_s"] = float(s.gstrf())
    s.download(); out["this_module_factor_check"] = float(s.factor_check())
    s.finalize()
    print(json.dumps(out), flush=True)
