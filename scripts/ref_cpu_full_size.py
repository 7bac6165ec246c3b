import sys, time, resource, json
sys.path.insert(0, '/root/repo')
from oracle.ref_run import run_reference
from pangulu_b200 import matrices
size, ranks = int(sys.argv[1]), int(sys.argv[2])
A = matrices.poisson3d(size)
t0 = time.time()
z = run_reference(A, 256, vtype="r64", perf=False, solve=False, blocks=False, nthread=1, ranks=ranks)
wall = time.time() - t0
ru = resource.getrusage(resource.RUSAGE_CHILDREN)
print(json.dumps({"size": size, "n": A.shape[0], "ranks": ranks, "t_gstrf": float(z["t_gstrf"]), "wall_total_s": wall,
                  "max_child_rss_gb": ru.ru_maxrss / 1e6, "keys": [k for k in z.keys() if k.startswith("t_")]}))
