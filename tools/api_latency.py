"""Wall time of the public Python calls at small and medium sizes (what a user of the reference's test-sized problems
sees): decomposition with a static permutation, decompose, solve, positive_semidefinite_matrix.  Best of 5 after a
warm-up call; MDB200_TIMING=1 adds the library's own stage breakdown on stderr."""
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


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


out = []
for n in [int(a) for a in sys.argv[1:]] or [100, 1000, 4000, 8192]:
    A, kw = workloads.shifted_goe(n, 5, 1.05)
    kw = dict(kw, permutation='decreasing_diagonal_values')
    S = A @ A.T / n + np.eye(n)
    b = np.ones(n)
    dec = psd.decomposition(A, **kw)
    row = dict(n=n,
               approximate_ms=1e3 * best(lambda: psd.decomposition(A, **kw)),
               approximate_and_read_L_ms=1e3 * best(lambda: psd.decomposition(A, **kw).L),
               psd_matrix_ms=1e3 * best(lambda: psd.positive_semidefinite_matrix(A, **kw)),
               decompose_ms=1e3 * best(lambda: matrix.decompose(S, permutation='decreasing_diagonal_values')),
               solve_resident_ms=1e3 * best(lambda: dec.solve(b)))
    out.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(REPO, 'gpurun_out', 'api_latency.json'), 'w'), indent=1)
