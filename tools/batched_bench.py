"""BASELINE config 5: many independent 512 x 512 matrix.approximate calls (per GPU: batch = 4096 / N).  Times the
device-resident batched factorization (matrices generated on the host once, uploaded, factored) and the same
through the batched C-ABI call with host buffers; reports matrices/s and GFLOP/s (n^3/3 each).

Under torchrun (one rank per GPU) every rank factors its own share -- independent units, no data-path
collective -- and rank 0 reports the whole-job throughput (total matrices / max over ranks of the device time):
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/batched_bench.py 512"""
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


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = 512
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    ctx = _native.context()          # device = LOCAL_RANK under torchrun
    lib = ctx.lib
    rng = np.random.default_rng(rank)
    mu = 1.05 * np.sqrt(2 * n)
    G = rng.standard_normal((batch, n, n))
    As = (G + G.transpose(0, 2, 1)) / 2 + mu * np.eye(n)[None]
    kw = workloads.f2_bounds(n)
    ps = np.argsort(-np.diagonal(As, axis1=1, axis2=2), axis=1).astype(np.int64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    P = _native.ptr
    out = dict(batch=batch, n=n)
    # device-resident: factor object with batch, set_matrix from host (untimed), run (timed with wall + sync)
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, batch, ctypes.byref(f)), 'create')
    ts = []
    for it in range(4):
        ctx.check(lib.mdb_factor_set_matrix(f, P(As), n, P(ps), _native.HOST), 'set')
        ctx.synchronize()
        l0 = ctx.launch_count()
        t = time.perf_counter()
        ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
        ctx.synchronize()
        ts.append(time.perf_counter() - t)
        out['launches_per_run'] = ctx.launch_count() - l0
    dt = min(ts[1:])
    out['device_s'] = dt
    out['device_matrices_per_s'] = batch / dt
    out['device_gflops'] = batch * n ** 3 / 3 / dt / 1e9
    ctx.set_profiling(True)
    ctx.check(lib.mdb_factor_set_matrix(f, P(As), n, P(ps), _native.HOST), 'set')
    ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
    ctx.synchronize()
    out['phases_ms'] = ctx.phase_ms()
    ctx.set_profiling(False)
    lib.mdb_factor_destroy(f)
    # through the batched C-ABI call with host buffers
    L = np.empty((batch, n, n))
    d, om, de = np.empty((batch, n)), np.empty((batch, n)), np.empty((batch, n))
    cl = np.zeros((batch, n), dtype=np.uint8)
    ts = []
    for it in range(3):
        t = time.perf_counter()
        ctx.check(lib.mdb_approximate_batched_f64(ctx.handle, batch, n, P(As), n, P(ps), ctypes.byref(bounds), P(L), n, P(d),
                                                  P(om), P(de), P(cl), _native.HOST), 'batched')
        ts.append(time.perf_counter() - t)
    dt = min(ts[1:])
    out['e2e_s'] = dt
    out['e2e_matrices_per_s'] = batch / dt
    out['n_clamped_mean'] = float(cl.sum(axis=1).mean())
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(ctx.device)
        dist.init_process_group('nccl', device_id=torch.device('cuda', ctx.device))
        t = torch.tensor([out['device_s'], out['e2e_s']], dtype=torch.float64, device=torch.device('cuda', ctx.device))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)      # reporting only: the work itself needs no exchange
        out.update(n_gpus=world, total_matrices=batch * world, device_s_max=float(t[0]), e2e_s_max=float(t[1]),
                   device_matrices_per_s_total=batch * world / float(t[0]),
                   e2e_matrices_per_s_total=batch * world / float(t[1]),
                   device_gflops_total=batch * world * n ** 3 / 3 / float(t[0]) / 1e9)
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    print(json.dumps(out))
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'batched_bench%s.json' % ('_%dgpu' % world if world > 1 else '')), 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
