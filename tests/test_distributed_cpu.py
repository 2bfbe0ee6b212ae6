"""The multi-GPU panel loop (matrix_decomposition_b200.distributed.factorize) on CPU: world_size 2 over
gloo with a numpy stand-in for the device steps.  Checks the protocol (ownership, slot rotation,
lookahead ordering) and that the distributed result equals the single-process blocked model / oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import workloads
from matrix_decomposition_b200 import distributed as md
from oracle import blocked_model as bm
from oracle import cpu_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, lookahead, agg, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from dist_numpy_steps import NumpySteps, GlooComm
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    A, kw = workloads.shifted_goe(n, 31, 1.05)
    p = orc.permutation_vector(A, kw['permutation'])
    Ap = A[p[:, None], p[None, :]]
    steps = NumpySteps(n, rank, world, Ap, min_diag_D=kw['min_diag_D'], max_diag_D=kw['max_diag_D'],
                       min_abs_value_D=kw['min_abs_value_D'], max_diag_B=kw['max_diag_B'], agg=agg)
    md.factorize(n, rank, world, steps, GlooComm(), lookahead=lookahead)
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), L=steps.local_L(), d=steps.d, omega=steps.omega,
             delta=steps.delta, clamped=steps.clamped, log=np.array(repr(steps.log)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('agg', [1, 2, 3])
@pytest.mark.parametrize('lookahead', [True, False])
def test_block_cyclic_loop_world2_gloo(tmp_path, lookahead, agg):
    n, world = 400, 2      # 4 column blocks, the last one partial; outer blocks of `agg` panels
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, lookahead, agg, str(tmp_path)), nprocs=world, join=True)
    A, kw = workloads.shifted_goe(n, 31, 1.05)
    ref = orc.approximate(A, **kw)
    kw2 = {k: v for k, v in kw.items() if k != 'permutation'}
    mod = bm.blocked_factor(A, ref['p'], nb=md.NB, **kw2)
    L = np.zeros((n, n))
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f'r{r}.npz'))
        # replicated vectors are identical on every rank and equal to the single-process model
        np.testing.assert_allclose(z['d'], mod['d'], rtol=1e-13)
        np.testing.assert_array_equal(z['clamped'], ref['clamped'])
        np.testing.assert_allclose(z['omega'], ref['omega'][ref['p']].astype(float), rtol=1e-10, atol=1e-14)
        for lb in range(md.local_blocks(n, r, world)):
            g0 = (lb * world + r) * md.NB
            w = min(md.NB, n - g0)
            L[:, g0:g0 + w] = z['L'][:, lb * md.NB: lb * md.NB + w]
        log = eval(str(z['log']))
        factored = [e[1] for e in log if e[0] == 'factor']
        assert factored == [k for k in range(4) if k % world == r]          # each rank factors only its own panels
        assert [e[1] for e in log if e[0] == 'absorb'] == [0, 1, 2, 3]      # every panel absorbed once, in order
        for e in log:
            if e[0] in ('factor', 'absorb'):
                assert e[2] == e[1] % 2                                      # message slot rotation
                assert e[3] == (e[1] // agg) * agg and e[4] == (e[1] // agg) % 2   # outer block / operand slot
            if e[0] == 'apply':
                assert e[3] >= e[1] and e[5] >= e[3] + e[4]                  # only columns beyond the panels used
    assert np.all(np.abs(L - mod['L']) <= 1e-12 * np.maximum(1.0, np.abs(mod['L'])))
    assert np.all(np.abs(L - ref['L']) <= 1e-10 * np.maximum(1.0, np.abs(ref['L'])))


def test_layout_helpers():
    assert md.local_blocks(400, 0, 2) == 2 and md.local_blocks(400, 1, 2) == 2
    assert md.local_blocks(300, 0, 2) == 2 and md.local_blocks(300, 1, 2) == 1
    assert md.local_cols(100, 1, 2) == 128          # a rank without blocks still gets a (dummy) slab
    assert md.panel_elems(400, 0) == 5 * 128 + 3 * 272 + 272 * 128
    assert md.panel_elems(400, 3) == 5 * 128
    from matrix_decomposition_b200 import _native
    lib = _native.load_library()
    for n, r, w in [(400, 0, 2), (400, 1, 2), (300, 1, 2), (100, 1, 2), (131072, 5, 8)]:
        assert lib.mdb_dist_local_cols(n, r, w) == md.local_cols(n, r, w)
