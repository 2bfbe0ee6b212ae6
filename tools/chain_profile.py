"""The sequential per-panel chain (k_diag_block -> k_panel -> inner update) in isolation: serialised phase times
(CUDA events per phase, mdb_ctx_set_profiling) of `decompose` (MDB_RULE_EXACT) and `approximate` (Reimer rule, F2)
at one size, and of the batched 512 x 512 mode.  Also the target of the ncu captures of k_diag_block / k_panel:
    ncu --set full --clock-control none --import-source on -k regex:'k_diag_block|k_panel' --launch-skip 40 \
        --launch-count 4 -o gpurun_out/chain python tools/chain_profile.py 8192 --quick"""
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import workloads  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402


def run(ctx, n, rule, A, p, bounds, batch=1, reps=3):
    lib = ctx.lib
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, batch, ctypes.byref(f)), 'create')
    out = {}
    ts = []
    for _ in range(reps):
        ctx.check(lib.mdb_factor_set_matrix(f, _native.ptr(A), n, _native.ptr(p), _native.HOST), 'set')
        ctx.synchronize()
        t = time.perf_counter()
        ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
        ctx.synchronize()
        ts.append(time.perf_counter() - t)
    out['ms'] = 1e3 * min(ts)
    out['tflops'] = batch * n ** 3 / 3 / min(ts) / 1e12
    ctx.set_profiling(True)
    ctx.check(lib.mdb_factor_set_matrix(f, _native.ptr(A), n, _native.ptr(p), _native.HOST), 'set')
    ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
    ctx.synchronize()
    out['phases_ms'] = ctx.phase_ms()
    ctx.set_profiling(False)
    lib.mdb_factor_destroy(f)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16384
    quick = '--quick' in sys.argv
    ctx = _native.context()
    res = dict(n=n)
    A, kw = workloads.shifted_goe(n, 5, 1.05)
    p = np.argsort(-A.diagonal()).astype(np.int64)
    eps = float(np.finfo(float).eps)
    b_reimer, k1 = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'], eps,
                                       None, kw['max_diag_B'])
    res['approximate'] = run(ctx, n, _native.RULE_REIMER, A, p, b_reimer, reps=1 if quick else 3)
    if not quick:
        S = workloads.lowrank_plus_diag(n, 3)
        ps = np.argsort(-S.diagonal()).astype(np.int64)
        b_exact, k2 = _native.make_bounds(_native.RULE_EXACT, 0.0, None, 0.0, 0.0)
        res['decompose'] = run(ctx, n, _native.RULE_EXACT, S, ps, b_exact)
        nb, batch = 512, 512
        rng = np.random.default_rng(0)
        G = rng.standard_normal((batch, nb, nb))
        mu = 1.05 * np.sqrt(2 * nb)
        As = (G + G.transpose(0, 2, 1)) / 2 + mu * np.eye(nb)[None]
        kwb = workloads.f2_bounds(nb)
        pb = np.argsort(-np.diagonal(As, axis1=1, axis2=2), axis=1).astype(np.int64)
        bb, k3 = _native.make_bounds(_native.RULE_REIMER, kwb['min_diag_D'], kwb['max_diag_D'], kwb['min_abs_value_D'], eps,
                                     None, kwb['max_diag_B'])
        res['batched_512x512'] = run(ctx, nb, _native.RULE_REIMER, As, pb, bb, batch=batch)
    print(json.dumps(res))
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'chain_profile_n%d.json' % n), 'w') as fh:
        json.dump(res, fh, indent=1)


if __name__ == '__main__':
    main()
