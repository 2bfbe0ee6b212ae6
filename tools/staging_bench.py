"""Pageable host -> device upload through the library's staged copy (staged_copy2d in mdb_api.cu): GB/s by number of
copy threads (MDB200_COPY_THREADS) and with / without streaming stores into the staging buffers (MDB200_COPY_NT)."""
import ctypes
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402

ctx = _native.context(0)
n = 32768
A = np.ones((n, n))          # 8.6 GB, pageable
dA = ctx.malloc(n * n * 8)
out = []
for nt in ('0', '1'):
    for T in (4, 8, 12, 16):
        os.environ['MDB200_COPY_THREADS'] = str(T)
        os.environ['MDB200_COPY_NT'] = nt
        best = 1e9
        for rep in range(3):
            t = time.perf_counter()
            ctx.check(ctx.lib.mdb_memcpy2d(ctx.handle, ctypes.c_void_p(dA), n, _native.ptr(A), n, n, n, _native.DEVICE,
                                           _native.HOST), 'mdb_memcpy2d')
            best = min(best, time.perf_counter() - t)
        out.append(dict(threads=T, streaming_stores=nt == '1', seconds=best, gbs=n * n * 8 / best / 1e9))
        print(out[-1], flush=True)
B = np.empty((n, n))
ctx.check(ctx.lib.mdb_memcpy2d(ctx.handle, _native.ptr(B), n, ctypes.c_void_p(dA), n, n, n, _native.HOST, _native.DEVICE), 'back')
print('round trip equal', bool(np.array_equal(A, B)))
print(json.dumps(out))
