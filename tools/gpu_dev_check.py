"""Step-by-step bring-up checks of libmdb200.so on a B200 (run through gpurun).  Prints one line per
check and keeps going after a failure, so a single GPU call yields as much information as possible.
The checker is the oracle / the numpy model of the blocked algorithm (test infrastructure)."""
import ctypes
import json
import os
import sys
import time
import traceback

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import workloads  # noqa: E402
import matrix_decomposition_b200 as matrix  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402
from oracle import cpu_oracle as orc  # noqa: E402
from oracle import blocked_model as bm  # noqa: E402

RESULTS = []


def report(name, ok, **info):
    RESULTS.append(dict(name=name, ok=bool(ok), **{k: (float(v) if isinstance(v, (np.floating, float)) else v)
                                                  for k, v in info.items()}))
    print(('PASS ' if ok else 'FAIL ') + name + ' ' + json.dumps(RESULTS[-1], default=str), flush=True)


def guarded(fn):
    def wrapper(*a, **k):
        try:
            fn(*a, **k)
        except Exception as e:  # keep going
            report(fn.__name__ + str(a), False, error=repr(e), tb=traceback.format_exc()[-600:])
    return wrapper


@guarded
def check_gemm(M, N, K, lower, beta1, neg, rg=0, cg=0):
    ctx = _native.context()
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    C0 = rng.standard_normal((M, N))
    C = C0.copy()
    ms = ctypes.c_double()
    rc = ctx.lib.mdb_debug_gemm_tn(ctx.handle, M, N, K, _native.ptr(A), K, _native.ptr(B), K, _native.ptr(C), N,
                                   int(lower), int(beta1), int(neg), rg, cg, 1, ctypes.byref(ms))
    ctx.check(rc, 'mdb_debug_gemm_tn')
    P = A @ B.T
    want = (C0 if beta1 else 0) + (-P if neg else P)
    if lower:
        mask = (rg + np.arange(M))[:, None] > (cg + np.arange(N))[None, :]
        want = np.where(mask, want, C0)
    err = np.abs(C - want).max()
    report(f'gemm M{M} N{N} K{K} lower{int(lower)} beta{int(beta1)} neg{int(neg)}', err < 1e-10 * K, err=err)


@guarded
def check_permute(n):
    A = workloads.goe(n, n)
    p = np.random.default_rng(n).permutation(n)
    B = matrix.permute.symmetric(A, p)
    report(f'permute n{n}', np.array_equal(B, A[p[:, None], p[None, :]]))


@guarded
def check_approx(name, A, kw, nb_note='', benign=True):
    t = time.time()
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    t_gpu = time.time() - t
    ref = orc.approximate(A, **kw)
    kw2 = {k: v for k, v in kw.items() if k != 'permutation'}
    mod = bm.blocked_factor(A, ref['p'], nb=128, **kw2)
    d = np.asarray(dec.d, dtype=float)
    L = dec.L
    d_ref = ref['d'].astype(float)
    B_ref = ref['L'] @ (d_ref[:, None] * ref['L'].T)
    B = L @ (d[:, None] * L.T)
    info = dict(
        n=A.shape[0], t_gpu=t_gpu,
        p_equal=bool(np.array_equal(dec.p, ref['p'])),
        clamped_equal=bool(np.array_equal(dec.clamped, ref['clamped'])),
        clamped_mismatch=int(np.count_nonzero(dec.clamped != ref['clamped'])),
        n_clamped=int(ref['clamped'].sum()),
        relB=float(np.linalg.norm(B - B_ref) / np.linalg.norm(B_ref)),
        maxdL_vs_oracle=float(np.max(np.abs(L - ref['L']) / np.maximum(1.0, np.abs(ref['L'])))),
        maxdL_vs_model=float(np.max(np.abs(L - mod['L']) / np.maximum(1.0, np.abs(mod['L'])))),
        maxdd_vs_oracle=float(np.max(np.abs(d - d_ref) / np.abs(d_ref).clip(1e-300))),
        maxdd_vs_model=float(np.max(np.abs(d - mod['d']) / np.abs(mod['d']).clip(1e-300))),
        first_bad_col=int(np.argmax(np.abs(d - mod['d']) > 1e-9 * np.abs(mod['d']))),
        domega=float(np.max(np.abs(np.asarray(dec.omega, dtype=float) - ref['omega'].astype(float)))),
        ddelta=float(np.max(np.abs(np.asarray(dec.delta, dtype=float) - ref['delta'].astype(float)))),
    )
    ok = info['p_equal'] and info['clamped_equal'] and info['relB'] < 1e-12
    if benign:
        ok = ok and info['maxdL_vs_oracle'] < 1e-10 and info['maxdd_vs_oracle'] < 1e-10
    report('approx ' + name, ok, **info)
    return dec, ref


@guarded
def check_solve(name, A, kw, nrhs):
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    n = A.shape[0]
    rng = np.random.default_rng(nrhs)
    b = rng.standard_normal((n, nrhs)) if nrhs > 1 else rng.standard_normal(n)
    x = dec.solve(b)
    x_ref = orc.solve(dec.L, np.asarray(dec.d, dtype=float), dec.p, b)
    err = np.max(np.abs(x - x_ref)) / np.max(np.abs(x_ref))
    y = dec.matrix_right_side_multiplication(b)
    y_ref = orc.multiply(dec.L, np.asarray(dec.d, dtype=float), dec.p, b)
    err_m = np.max(np.abs(y - y_ref)) / np.max(np.abs(y_ref))
    report(f'solve {name} nrhs{nrhs}', err < 1e-10 and err_m < 1e-12, solve_err=err, multiply_err=err_m)


@guarded
def check_decompose(n, seed, rt, perm):
    A = workloads.lowrank_plus_diag(n, seed)
    dec = matrix.decompose(A, permutation=perm, return_type=rt)
    ref = orc.decompose(A, perm, 'LL' if rt == 'LL' else 'LDL')
    if rt == 'LL':
        err = np.max(np.abs(dec.L - ref['L']))
    elif rt == 'LDL':
        err = max(np.max(np.abs(dec.L - ref['L'])), np.max(np.abs(dec.d - ref['d']) / ref['d']))
    else:
        LD = ref['L'].copy()
        np.fill_diagonal(LD, ref['d'])
        err = np.max(np.abs(dec.LD - LD))
    b = np.random.default_rng(1).standard_normal((n, 3))
    x = dec.solve(b)
    res = np.max(np.abs(A @ x - b)) / np.max(np.abs(b))
    report(f'decompose n{n} {rt} {perm}', err < 1e-10 and res < 1e-9 and np.array_equal(dec.p, ref['p']), err=err,
           residual=res)


@guarded
def check_decompose_failure():
    A = workloads.reference_fixture(40)
    _, info = orc.potrf_lower(A)
    try:
        matrix.decompose(A)
        report('decompose failure index', False, note='no exception')
    except matrix.errors.NoDecompositionPossibleWithProblematicSubdecompositionError as e:
        report('decompose failure index', e.problematic_leading_principal_submatrix_index == info - 1,
               got=int(e.problematic_leading_principal_submatrix_index), want=info - 1)


@guarded
def timing(n, s=1.05, exact=False):
    """HBM-resident factorization timing with the device generator (family F2)."""
    ctx = _native.context()
    lib = ctx.lib
    mu = s * np.sqrt(2 * n)
    diag = np.empty(n)
    ctx.check(lib.mdb_generate_diag_f2(ctx.handle, n, 1234, mu, _native.ptr(diag)), 'diag')
    p = np.argsort(-diag).astype(np.int64)
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'create')
    kw = workloads.f2_bounds(n, s)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       np.finfo(float).eps, None, kw['max_diag_B'])
    if exact:
        bounds, keep = _native.make_bounds(_native.RULE_EXACT, 0.0, None, 0.0, 0.0)
    out = {}
    for rep in range(3):
        ctx.check(lib.mdb_factor_generate_f2(f, 1234, mu, _native.ptr(p)), 'gen')
        ctx.synchronize()
        ctx.set_profiling(rep == 2)
        t = time.perf_counter()
        ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
        ctx.synchronize()
        dt = time.perf_counter() - t
        out['profiled' if rep == 2 else 'plain'] = dt
        if rep == 2:
            out['phases_ms'] = ctx.phase_ms()
    ctx.set_profiling(False)
    res = ctypes.c_double()
    ctx.check(lib.mdb_factor_identity_residual(f, 2, 99, ctypes.byref(res)), 'identity')
    clamped = np.zeros(n, dtype=np.uint8)
    d = np.empty(n)
    ctx.check(lib.mdb_factor_get(f, _native.TYPE_LDL, None, 0, _native.HOST, _native.ptr(d), None, None,
                                 _native.ptr(clamped), None), 'get')
    lib.mdb_factor_destroy(f)
    gflops = n ** 3 / 3 / out['plain'] / 1e9
    report(f'timing n{n} exact{int(exact)} lookahead{os.environ.get("MDB200_LOOKAHEAD", "1")}', res.value < 1e-10, seconds=out['plain'], gflops=gflops, frac_of_37_1=gflops / 37100,
           phases_ms=out.get('phases_ms'), identity_residual=res.value, n_clamped=int(clamped.sum()),
           d_min=float(d.min()), d_max=float(d.max()))


def main():
    quick = '--quick' in sys.argv
    print('device context:', _native.context().device, flush=True)
    if '--timing-only' in sys.argv:
        sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [4096, 16384]
        for n in sizes:
            timing(n)
            timing(n, exact=True)
        return 0
    # 1. the tile kernel
    check_gemm(128, 64, 16, False, False, False)
    check_gemm(128, 64, 128, False, True, False)
    check_gemm(256, 192, 128, False, True, True)
    check_gemm(300, 300, 128, True, True, False, 128, 128)
    check_gemm(1000, 1000, 128, True, True, False, 256, 256)
    check_gemm(77, 50, 32, False, False, False)
    check_gemm(5, 700, 128, False, True, True)
    # 2. gather
    check_permute(10)
    check_permute(333)
    # 3. factorization against the oracle and the numpy model
    A2 = np.array([[2.0, 1.0], [1.0, -3.0]])
    check_approx('known2x2', A2, dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values'), benign=True)
    for n in (10, 64, 128, 129, 200, 300, 700):
        A, kw = workloads.shifted_goe(n, n, 1.05)
        check_approx(f'f2 n{n} s1.05', A, kw)
    A, kw = workloads.shifted_goe(400, 3, 0.9)
    check_approx('f2 n400 s0.9 (explosive)', A, kw, benign=False)
    check_approx('goe n300 1e-6 (explosive)', workloads.goe(300, 4),
                 dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values'), benign=False)
    check_approx('fixture96 vecB', workloads.reference_fixture(96),
                 dict(min_diag_D=None, min_diag_B=np.arange(96) / 96.0, max_diag_B=np.arange(96) + 30.0,
                      permutation='none'), benign=False)
    check_approx('fixture10 grid-like', workloads.reference_fixture(10),
                 dict(min_diag_D=0, max_diag_D=1, max_diag_B=1, permutation=workloads.reference_permutation(10)),
                 benign=False)
    # 4. solve / multiply
    A, kw = workloads.shifted_goe(300, 5, 1.05)
    for nrhs in (1, 3, 40):
        check_solve('f2 n300', A, kw, nrhs)
    A, kw = workloads.shifted_goe(1000, 6, 1.05)
    check_solve('f2 n1000', A, kw, 1)
    check_solve('f2 n1000', A, kw, 64)
    # 5. decompose
    for rt in ('LL', 'LDL', 'LDL_compressed'):
        check_decompose(200, 2, rt, 'decreasing_diagonal_values')
    check_decompose(513, 3, 'LDL', 'none')
    check_decompose_failure()
    # 6. timings
    if not quick:
        for n in (4096, 8192, 16384):
            timing(n)
    n_fail = sum(not r['ok'] for r in RESULTS)
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'dev_check.json'), 'w') as fh:
        json.dump(RESULTS, fh, indent=1, default=str)
    print(f'{len(RESULTS) - n_fail} passed, {n_fail} failed')
    return 0  # the script itself always exits 0; read the lines


if __name__ == '__main__':
    sys.exit(main())
