"""Outer-block schedule of the factorization at config-2 size: device time of `decompose` (rule EXACT) at n = 16384 for
a few threshold triples (remaining rows from which 2 / 4 / 8 panels are aggregated, mdb_ctx_set_blocking)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import workloads  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = _native.context(0)
stream = torch.cuda.Stream(priority=-1)
ctx.set_stream(stream.cuda_stream)
lib = ctx.lib
A = workloads.lowrank_plus_diag(n, 3)
p = np.argsort(-A.diagonal(), kind='stable').astype(np.int64)
bounds, keep = _native.make_bounds(_native.RULE_EXACT, 0.0, None, 0.0, 0.0)
out = []
for t2, t4, t8 in ((4096, 8192, 32768), (3072, 8192, 32768), (2048, 8192, 32768), (6144, 8192, 32768), (4096, 6144, 32768),
                   (4096, 12288, 32768), (3072, 6144, 32768), (4096, 8192, 16384), (4096, 8192, 12288), (2048, 6144, 12288),
                   (4096, 8192, 32768)):
    ctx.set_blocking(0, t2, t4, t8)
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'create')
    ms = bench._event_ms(stream, lambda: ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run'), reps=4,
                         before=lambda: ctx.check(lib.mdb_factor_set_matrix(f, _native.ptr(A), n, _native.ptr(p), _native.HOST), 'set'))
    lib.mdb_factor_destroy(f)
    out.append(dict(t2=t2, t4=t4, t8=t8, ms=ms, frac=n ** 3 / 3 / (ms * 1e-3) / 1e12 / bench.FP64_DMMA_PEAK_TFLOPS))
    print(json.dumps(out[-1]), flush=True)
ctx.set_stream(None)
json.dump(out, open(os.path.join(REPO, 'gpurun_out', 'blocking_sweep_n%d.json' % n), 'w'), indent=1)
