"""Stage breakdown of the one-shot C-ABI call with pinned host buffers (the bench's e2e leg): MDB200_TIMING=1 makes
libmdb200 print create / set_matrix / run / get / destroy wall times to stderr."""
import argparse
import math
import os
import sys

os.environ['MDB200_TIMING'] = '1'
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--size', type=int, default=65536)
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--shift', type=float, default=None, help='s of the F2 family (default: the bench value 1.05; 0.9 = indefinite, many rows with omega < 1)')
a = ap.parse_args()
if a.shift is not None:
    bench.S_SHIFT = a.shift
n = a.size
ctx = _native.context(0)
mu = bench.S_SHIFT * math.sqrt(2 * n)
kw = bench.f2_bounds(n)
bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                   float(np.finfo(np.float64).eps), None, kw['max_diag_B'])
seed = 20261019
diag = np.empty(n)
ctx.check(ctx.lib.mdb_generate_diag_f2(ctx.handle, n, seed, mu, _native.ptr(diag)), 'mdb_generate_diag_f2')
p = np.argsort(-diag).astype(np.int64)
print(bench.run_e2e(ctx, n, seed, mu, p, bounds, a))
