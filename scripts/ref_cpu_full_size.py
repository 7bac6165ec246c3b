#!/usr/bin/env python
"""The UNMODIFIED reference CPU platform (oracle/_ref) on a FULL BASELINE configuration, once -- the same-configuration
baseline the bench's bounded `--impl reference` arm cannot afford per step.

    python scripts/ref_cpu_full_size.py poisson3d_100 8        # workload name of bench.py, reference ranks
    python scripts/ref_cpu_full_size.py 100 8                  # (a bare number = poisson3d of that size)

Prints one JSON line: the reference's own barrier-to-barrier gstrf interval (max over the ranks), the whole run including its
pangulu_init, and the largest rank's resident set.  Results of this round: profiles/r02_reference_cpu_full_*.json."""
import json
import os
import resource
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.ref_run import run_reference  # noqa: E402

name, ranks = sys.argv[1], int(sys.argv[2])
if name.isdigit():
    gen, vtype, size = "poisson3d", "r64", int(name)
else:
    gen, vtype, size, _ = bench.WORKLOADS[name]
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 256
A = bench.make_matrix(gen, size, vtype)
t0 = time.time()
z = run_reference(A, nb, vtype=vtype, perf=False, solve=False, blocks=False, nthread=1, ranks=ranks)
wall = time.time() - t0
ru = resource.getrusage(resource.RUSAGE_CHILDREN)
print(json.dumps({"workload": bench.workload_name(gen, size, vtype, nb), "n": int(A.shape[0]), "nnz_A": int(A.nnz), "ranks": ranks,
                  "t_gstrf_s": float(z["t_gstrf"]), "wall_total_s_including_pangulu_init": wall, "max_rank_rss_gb": ru.ru_maxrss / 1e6,
                  "host_cpus": os.cpu_count()}))
