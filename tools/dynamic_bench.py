"""Dynamic pivoting ('maximal_stability', the reference's default for permutation=None, Reimer.py:415-416):
wall time of approximation.positive_semidefinite.decomposition through the public API at a few sizes."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import workloads  # noqa: E402
import matrix_decomposition_b200 as matrix  # noqa: E402

out = []
for n in [int(a) for a in sys.argv[1:]] or [1000, 2000, 4000, 8000]:
    A, kw = workloads.shifted_goe(n, 5, 1.05)
    kw = dict(kw, permutation='maximal_stability')
    matrix.approximation.positive_semidefinite.decomposition(A[:256, :256].copy(), **dict(workloads.f2_bounds(256), permutation='maximal_stability'))
    dt = 1e30
    for _ in range(3):    # best of three: the first call at a size pays one-time allocations (pinned staging, buffers)
        t = time.perf_counter()
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        dt = min(dt, time.perf_counter() - t)
    d = np.asarray(dec.d, dtype=float)
    out.append(dict(n=n, seconds=dt, gflops=n ** 3 / 3 / dt / 1e9, n_clamped=int(dec.clamped.sum()), d_min=float(d.min())))
    print(json.dumps(out[-1]), flush=True)
os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(REPO, 'gpurun_out', 'dynamic_bench.json'), 'w'), indent=1)
