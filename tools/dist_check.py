"""Multi-GPU bring-up (torchrun, one rank per GPU): the block-cyclic NCCL path against the single-GPU path
and the oracle at moderate n, then a timing + identity check at a larger n."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import workloads  # noqa: E402
import matrix_decomposition_b200 as matrix  # noqa: E402
from matrix_decomposition_b200 import distributed as md  # noqa: E402


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', str(rank)))
    torch.cuda.set_device(local)
    md.init_process_group(local)
    out = {}
    from matrix_decomposition_b200 import _native
    for n, fixed in ((700, 1), (1500, 2), (1500, 4)):
        _native.context(local).set_blocking(fixed=fixed)
        A, kw = workloads.shifted_goe(n, 3, 1.05)
        p = np.argsort(-A.diagonal()).astype(np.int64)
        Ap = A[p[:, None], p[None, :]]
        bk = dict(min_diag_D=kw['min_diag_D'], max_diag_D=kw['max_diag_D'], min_abs_value_D=kw['min_abs_value_D'],
                  min_abs_value_L=float(np.finfo(float).eps), max_diag_B=kw['max_diag_B'])
        for la in (True, False):
            steps, d, omega, delta, clamped = md.approximate_distributed(Ap, n, bk, p=p, lookahead=la)
            ident = md.identity_residual(steps, d, omega, delta)
            Lloc = steps.S.cpu().numpy()
            # the consumers of the distributed factor: solve on the slabs, gather to rank 0 (collective calls)
            fac = md.DistributedFactor(steps, d, p, omega, delta, clamped)
            b = np.random.default_rng(n).standard_normal((n, 3))
            x = fac.solve(b)
            x1 = fac.solve(b[:, 0])
            gathered = fac.gather(0)
            if rank == 0:
                dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
                ok = np.array_equal(clamped, dec.clamped) and np.array_equal(d, np.asarray(dec.d, dtype=float))
                for lb in range(md.local_blocks(n, 0, world)):
                    g0 = (lb * world) * md.NB
                    w = min(md.NB, n - g0)
                    ok = ok and np.array_equal(Lloc[:, lb * md.NB: lb * md.NB + w], dec.L[:, g0:g0 + w])
                x_ref = dec.solve(b)
                out[f'n{n}_agg{fixed}_lookahead{int(la)}'] = dict(
                    bitwise_equal_single_gpu=bool(ok), identity=ident,
                    gather_bitwise_equal=bool(np.array_equal(gathered.L, dec.L)
                                              and np.array_equal(np.asarray(gathered.omega, dtype=float), np.asarray(dec.omega, dtype=float))),
                    gathered_solve_equal=bool(np.array_equal(gathered.solve(b), x_ref)),
                    dist_solve_rel_err=float(np.max(np.abs(x - x_ref)) / np.max(np.abs(x_ref))),
                    dist_solve_1rhs_rel_err=float(np.max(np.abs(x1 - x_ref[:, 0])) / np.max(np.abs(x_ref[:, 0]))))
                del gathered
            steps.close()
            del steps
    _native.context(local).set_blocking()
    big = int(os.environ.get('DIST_N', '32768'))
    t = time.time()
    res = md.bench_f2(big, 1.05, steps=1, warmup=1, seed=7)
    if rank == 0:
        res['gflops'] = big ** 3 / 3 / (res['ms_per_step'] * 1e-3) / 1e9
        res['wall'] = time.time() - t
        out[f'bench_n{big}_gpus{world}'] = res
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(REPO, 'gpurun_out', f'dist_check_{world}gpu.json'), 'w') as fh:
            json.dump(out, fh, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
