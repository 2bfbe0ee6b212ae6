"""Efficiency of the DMMA tile kernel vs K and beta (guides the optimisation of the trailing update)."""
import ctypes, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from matrix_decomposition_b200 import _native

ctx = _native.context()
lib = ctx.lib
rng = np.random.default_rng(0)
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
C = rng.standard_normal((M, N))
out = []
for K in (128, 256, 512):
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    for (lower, beta1) in ((1, 1), (0, 1), (0, 0)):
        ms = ctypes.c_double()
        Cc = C.copy()
        P = _native.ptr
        ctx.check(lib.mdb_debug_gemm_tn(ctx.handle, M, N, K, P(A), K, P(B), K, P(Cc), N, lower, beta1, 0, 0, 0, 10,
                                        ctypes.byref(ms)), 'gemm')
        flop = (M * N if lower else 2 * M * N) * K
        r = dict(M=M, K=K, lower=lower, beta1=beta1, ms=ms.value, tflops=flop / ms.value / 1e9, frac=flop / ms.value / 1e9 / 37.1)
        out.append(r)
        print(json.dumps(r), flush=True)
json.dump(out, open(os.path.join(REPO, 'gpurun_out', 'gemm_experiments.json'), 'w'), indent=1)
