import sys, time
t0 = time.time()
sys.path[:0] = ['tests', '.']
import test_zz_registered_gpu as t
from pangulu_b200 import matrices
t.test_reference_runs_on_the_registered_platform('a', lambda: matrices.poisson3d(12), 32)
print('poisson3d-12-nb32 on the registered platform: PASS', round(time.time() - t0, 1), 's', flush=True)
t.test_reference_runs_on_the_registered_platform('b', lambda: matrices.stencil27(8), 128)
print('stencil27-8-nb128 on the registered platform: PASS', round(time.time() - t0, 1), 's', flush=True)
