"""Run-to-run and path-to-path determinism of the factorization (lookahead races would show up here) and of dynamic
pivoting (last-block merge of the arg-min)."""
import ctypes
import json
import math
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import workloads  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402


def run_resident(ctx, n, seed, p, bounds, lookahead, want_L=False):
    lib = ctx.lib
    mu = 1.05 * math.sqrt(2 * n)
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'create')
    ctx.check(lib.mdb_factor_generate_f2(f, seed, mu, _native.ptr(p)), 'gen')
    ctx.set_profiling(not lookahead)   # profiling mode serialises everything on one stream
    ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
    ctx.set_profiling(False)
    d = np.empty(n)
    cl = np.zeros(n, dtype=np.uint8)
    L = np.empty((n, n)) if want_L else None
    ctx.check(lib.mdb_factor_get(f, _native.TYPE_LDL, _native.ptr(L), n, _native.HOST, _native.ptr(d), None, None,
                                 _native.ptr(cl), None), 'get')
    lib.mdb_factor_destroy(f)
    return d, cl, L


def main():
    ctx = _native.context()
    lib = ctx.lib
    out = {}
    for n in [int(a) for a in sys.argv[1:]] or [8192, 20000]:
        mu = 1.05 * math.sqrt(2 * n)
        seed = 99
        diag = np.empty(n)
        ctx.check(lib.mdb_generate_diag_f2(ctx.handle, n, seed, mu, _native.ptr(diag)), 'diag')
        p = np.argsort(-diag).astype(np.int64)
        kw = workloads.f2_bounds(n)
        bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                           float(np.finfo(float).eps), None, kw['max_diag_B'])
        want_L = n <= 8192
        d0, c0, L0 = run_resident(ctx, n, seed, p, bounds, False, want_L)     # serialised reference run
        res = dict(n_clamped_serial=int(c0.sum()))
        for rep in range(3):
            d1, c1, L1 = run_resident(ctx, n, seed, p, bounds, True, want_L)
            res[f'lookahead_run{rep}'] = dict(d_equal=bool(np.array_equal(d0, d1)), clamped_equal=bool(np.array_equal(c0, c1)),
                                             n_clamped=int(c1.sum()), L_equal=(bool(np.array_equal(L0, L1)) if want_L else None),
                                             max_rel_d=float(np.max(np.abs(d1 - d0) / np.abs(d0))))
        # host path: unpermuted matrix on the host, gathered by the library
        if n <= 20000 or os.environ.get('DET_HOST_ALWAYS'):
            dA = ctx.malloc(n * n * 8)
            ctx.check(lib.mdb_generate_f2(ctx.handle, n, seed, mu, ctypes.c_void_p(dA), n), 'gen2')
            A = np.empty((n, n))
            ctx.check(lib.mdb_memcpy(ctx.handle, _native.ptr(A), ctypes.c_void_p(dA), n * n * 8, _native.HOST, _native.DEVICE), 'cp')
            ctx.free(dA)
            L = np.empty((n, n)); d = np.empty(n); om = np.empty(n); de = np.empty(n); cl = np.zeros(n, dtype=np.uint8)
            P = _native.ptr
            ctx.check(lib.mdb_approximate_f64(ctx.handle, n, P(A), n, P(p), ctypes.byref(bounds), _native.TYPE_LDL, P(L), n,
                                              P(d), P(om), P(de), P(cl), _native.HOST, None), 'approx')
            res['host_path'] = dict(d_equal=bool(np.array_equal(d0, d)), clamped_equal=bool(np.array_equal(c0, cl)),
                                    n_clamped=int(cl.sum()), diag_equal=bool(np.array_equal(A.diagonal(), diag)),
                                    sym=bool(np.array_equal(A, A.T)),
                                    first_diff=int(np.argmax(d0 != d)) if not np.array_equal(d0, d) else -1,
                                    n_diff=int(np.count_nonzero(d0 != d)))
        out[f'n{n}'] = res
        print(json.dumps({f'n{n}': res}), flush=True)
    # dynamic pivoting: the arg-min is merged by whichever block finishes last -- p, d and L must not depend on that
    import matrix_decomposition_b200 as matrix
    nd = 6000
    A, kw = workloads.shifted_goe(nd, 7, 0.97)
    runs = []
    for rep in range(3):
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **dict(kw, permutation='maximal_stability'))
        runs.append((np.asarray(dec.p).copy(), np.asarray(dec.d, dtype=float).copy(), np.array(dec.L)))
        del dec
    res = dict(p_equal=all(np.array_equal(runs[0][0], r[0]) for r in runs[1:]),
               d_equal=all(np.array_equal(runs[0][1], r[1]) for r in runs[1:]),
               L_equal=all(np.array_equal(runs[0][2], r[2]) for r in runs[1:]))
    out[f'dynamic_n{nd}'] = res
    print(json.dumps({f'dynamic_n{nd}': res}), flush=True)
    with open(os.path.join(REPO, 'gpurun_out', 'determinism.json'), 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
