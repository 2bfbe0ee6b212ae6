"""Correctness stress of the tile kernels at large sizes (races show up as rare wrong tiles)."""
import ctypes, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from matrix_decomposition_b200 import _native
ctx = _native.context(); lib = ctx.lib
rng = np.random.default_rng(1)
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for K in (128,):
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C0 = rng.standard_normal((M, N))
    want_full = C0 + A @ B.T
    mask = np.arange(M)[:, None] > np.arange(N)[None, :]
    want = np.where(mask, want_full, C0)
    for rep in range(4):
        C = C0.copy(); ms = ctypes.c_double(); P = _native.ptr
        ctx.check(lib.mdb_debug_gemm_tn(ctx.handle, M, N, K, P(A), K, P(B), K, P(C), N, 1, 1, 0, 0, 0, 1, ctypes.byref(ms)), 'g')
        err = np.abs(C - want)
        bad = np.argwhere(err > 1e-9)
        print(json.dumps(dict(variant=os.environ.get('MDB200_TMA_VARIANT', '0'), tma=os.environ.get('MDB200_TMA', '1'), K=K, rep=rep,
                              max_err=float(err.max()), n_bad=int(len(bad)),
                              bad_tiles=sorted({(int(i) // 128, int(j) // 64) for i, j in bad[:2000]})[:12])), flush=True)
