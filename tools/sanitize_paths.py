"""Small invocations of every kernel family of libmdb200.so in one process (a few seconds): written as the target of
`compute-sanitizer --tool memcheck python tools/sanitize_paths.py` (`--tool racecheck` with SANITIZE_SMALL=1).  The GPU
pool of this project refuses compute-sanitizer, so here it only serves as a fast everything-launches-once pass."""
import ctypes
import os
import sys

os.environ.setdefault('MDB200_TAIL_MIN', '1024')
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import workloads  # noqa: E402
import matrix_decomposition_b200 as matrix  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402

small = bool(os.environ.get('SANITIZE_SMALL'))
psd = matrix.approximation.positive_semidefinite


def step(name):
    print('...', name, flush=True)


step('static approximate, two rules of the diagonal-block kernel')
for n, s in ((300, 1.05), (700, 0.9)):
    A, kw = workloads.shifted_goe(n, 3, s)
    dec = psd.decomposition(A, **kw)
    assert np.all(np.isfinite(np.asarray(dec.d, dtype=float)))
step('decompose + solve (1 and 16 right-hand sides) + multiply')
n = 600
A, kw = workloads.shifted_goe(n, 4, 1.05)
dec = psd.decomposition(A, **kw)
b = np.random.default_rng(0).standard_normal((n, 16))
x1 = dec.solve(b[:, 0])
x16 = dec.solve(b)
assert np.allclose(x16[:, 0], x1)
step('symmetric permutation')
p = np.random.default_rng(1).permutation(n)
assert np.array_equal(matrix.permute.symmetric(A, p), A[np.ix_(p, p)])
step('batched small matrices')
As = np.stack([workloads.shifted_goe(200, 10 + k, 1.05)[0] for k in range(4)])
res = psd.decomposition_batched(As, **workloads.f2_bounds(200))
assert np.all(np.isfinite(res.d))
step('dynamic pivoting (graph replay from n = 384)')
for n in ((200,) if small else (200, 450)):
    A, kw = workloads.shifted_goe(n, 5, 1.05)
    for method in ('maximal_stability', 'minimal_difference'):
        dec = psd.decomposition(A, **dict(kw, permutation=method))
        assert sorted(np.asarray(dec.p).tolist()) == list(range(n))
step('GMW / SE families')
g = np.random.default_rng(2).standard_normal((150, 150))
for fn in ('GMW_81', 'GMW_T2', 'SE_90', 'SE_99', 'SE_T1'):
    B = getattr(psd, fn)((g + g.T) / 2)
    assert np.all(np.isfinite(B))
if not small:
    step('one-shot call with a pinned L streamed out, early shipment of the late rows forced on')
    ctx = _native.context()
    n = 4300
    A, kw = workloads.shifted_goe(n, 78, 0.9)
    p = np.argsort(-A.diagonal()).astype(np.int64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    L = ctx.pinned_array((n, n))
    d, om, de = np.empty(n), np.empty(n), np.empty(n)
    cl = np.zeros(n, dtype=np.uint8)
    P = _native.ptr
    ctx.check(ctx.lib.mdb_approximate_f64(ctx.handle, n, P(A), n, P(p), ctypes.byref(bounds), _native.TYPE_CODES['LDL'],
                                          P(L), n, P(d), P(om), P(de), P(cl), _native.HOST, None), 'one-shot')
    assert np.all(np.isfinite(L))
    _native.free_pinned(L)
print('sanitize_paths: all steps ran')
