"""Symmetric permutation B[i, j] = A[p[i], p[j]] (dense/permute.py:26) on the device: HBM GB/s of the algorithmic
16 n^2 bytes (read + write of the full matrix) for a random p and for the identity, device buffers in and out."""
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from matrix_decomposition_b200 import _native  # noqa: E402

HBM_PEAK_GBS = 6547.5


def main():
    ctx = _native.context()
    lib = ctx.lib
    out = []
    for n in (16384, 32768, 65536):
        dA, dB = ctx.malloc(n * n * 8), ctx.malloc(n * n * 8)
        ctx.check(lib.mdb_generate_f2(ctx.handle, n, 1, 10.0, ctypes.c_void_p(dA), n), 'gen')
        for name, p in (('random', np.random.default_rng(0).permutation(n).astype(np.int64)),
                        ('identity', np.arange(n, dtype=np.int64))):
            ts = []
            for _ in range(4):
                ctx.synchronize()
                t = time.perf_counter()
                ctx.check(lib.mdb_permute_sym_f64(ctx.handle, n, ctypes.c_void_p(dA), n, _native.ptr(p), ctypes.c_void_p(dB), n,
                                                  _native.DEVICE), 'permute')
                ctx.synchronize()
                ts.append(time.perf_counter() - t)
            dt = min(ts[1:])
            out.append(dict(n=n, p=name, seconds=dt, gbs=16.0 * n * n / dt / 1e9, frac_hbm=16.0 * n * n / dt / 1e9 / HBM_PEAK_GBS))
            print(json.dumps(out[-1]), flush=True)
        ctx.free(dA)
        ctx.free(dB)
    with open(os.path.join(REPO, 'gpurun_out', 'permute_bench.json'), 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
