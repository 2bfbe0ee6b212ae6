"""GMW_81 / GMW_T1 (approximation/positive_semidefinite/GMW_SE.py of the reference) through the public API: wall time at a
few sizes (best of three).  The factorization is unblocked by construction -- d_k depends on the whole column below the
pivot -- so it is a sequence of rank-1 updates: 8 (n - k)^2 bytes of HBM traffic and three launches per column."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import workloads  # noqa: E402
import matrix_decomposition_b200 as matrix  # noqa: E402

psd = matrix.approximation.positive_semidefinite
out = []
for n in [int(a) for a in sys.argv[1:]] or [1000, 2000, 4000]:
    A = workloads.goe(n, 3) + 0.5 * np.sqrt(2 * n) * np.eye(n)     # indefinite: both phases are exercised
    for name in ('GMW_81', 'GMW_T1'):
        fn = getattr(psd, name)
        dt = 1e30
        for _ in range(3):
            t = time.perf_counter()
            B = fn(A, min_diag_D=1e-3)
            dt = min(dt, time.perf_counter() - t)
        out.append(dict(algorithm=name, n=n, seconds=dt, us_per_column=1e6 * dt / n,
                        update_gbs=16.0 * n ** 3 / 3 / dt / 1e9, finite=bool(np.isfinite(B).all())))
        print(json.dumps(out[-1]), flush=True)
os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(REPO, 'gpurun_out', 'gmw_bench.json'), 'w'), indent=1)
