#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native matrix-decomposition hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size SIZE]

A "step" is one modified LDL^T factorization (matrix.approximate /
approximation.positive_semidefinite.decomposition) of one synthetic dense symmetric float64 matrix of
family F2 (GOE + mu I, mu = 1.05 sqrt(2n), bounds min_diag_D = min_abs_value_D = 0.015 mu,
max_diag_D = 0.9 mu, max_diag_B = 2 mu, permutation 'decreasing_diagonal_values').

  N = 1 : n = 65536 (34 GB), BASELINE.json configs[2].  `value` = n^3/3 flop / device time of the whole job
          with the (unpermuted) matrix already resident in HBM: symmetric permutation gather + factorization +
          finalisation of L (omega scaling / thresholding); `e2e` = the same metric through the C-ABI one-shot
          call mdb_approximate_f64 with pinned HOST buffers (H2D of A, D2H of L, d, omega, delta, clamped inside
          the timed region); `e2e_numpy` = the public Python call decomposition(A) on a pageable numpy array.
          `parity` = the same code path on the n = 6144 matrix of the CPU sample against the oracle's result
          (p, clamped set, d, L, B).  `extra` = the other BASELINE configs that fit one GPU: config 2
          (decompose n = 16384 + solve with 1 and 1024 right-hand sides), the symmetric permutation at
          n = 65536, config 5's per-GPU share (512 matrices of 512 x 512, batched).
  N > 1 : n = 131072 (137 GB), 1-D block-cyclic columns over N ranks (torchrun), NCCL panel broadcast
          (BASELINE.json configs[3]); fixed total work, so "scaling": "strong" between N = 2, 4, 8.
  --impl reference : the CPU arm.  The reference is pure Python and cannot travel to the GPU box; its
          restatement (oracle/mdo_oracle.c, pinned against the reference's golden vectors) is timed on all
          host cores on a bounded sample (n = 6144 of the same family) and reported in the same unit.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'approximate FP64 GFLOP/s (n^3/3)'
UNIT = 'GFLOP/s'
FP64_DMMA_PEAK_TFLOPS = 37.1   # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak.jsonl);
                               # MEASURED_PEAKS.json holds only HBM and bf16 peaks
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the bulk kernel from one `ncu --set full` capture
# (round 2, the current kernel: profiles/r02_ncu_summary.txt)
NCU_TRAFFIC = 46.04e9   # bulk update part (b), grid 234240 tiles, 109.3 ms: read 30.96 GB + write 15.08 GB
NCU_TRAFFIC_NOTE = ('ncu --set full capture of the LARGEST bulk launch (part (b) of the first outer block, grid '
                    '234240 tiles, 109.3 ms, DMMA pipe 95.4 % active): dram read 30.96 GB + write 15.08 GB vs '
                    'algorithmic 31.7 GB (C triangle read + written 30.7 GB, operands 1.0 GB); the excess are '
                    'operand re-reads that missed L2 (hit rate 87 %; profiles/r02_ncu_summary.txt)')
CPU_SAMPLE_N = 6144
S_SHIFT = 1.05


def f2_bounds(n):
    import workloads
    return workloads.f2_bounds(n, S_SHIFT)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        fd, self.path = tempfile.mkstemp(suffix='.csv')
        os.close(fd)
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            busy = [c for c in sm if c > 0.5 * max(sm)] or sm
            out['sm_mhz'] = float(np.median(busy))
            out['sm_max_mhz'] = float(max(smax))
        out['reasons'] = sorted(reasons)
        return out


def cpu_sample_matrix(n=CPU_SAMPLE_N):
    import workloads
    return workloads.shifted_goe(n, 4242, S_SHIFT)


def cpu_baseline(n=CPU_SAMPLE_N, reps=1, keep_result=False):
    """The oracle port of the reference's column loop, all host cores (OpenMP), bounded sample.
    Returns (baseline dict, seconds, oracle result or None)."""
    from oracle import cpu_oracle as orc
    A, kw = cpu_sample_matrix(n)
    # all host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    cores = orc.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    orc.approximate(A[:256, :256].copy(), **f2_bounds(256))  # build + warm the library
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        res = orc.approximate(A, **kw)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    base = dict(value=n ** 3 / 3 / best / 1e9, unit=UNIT, cores=cores, kind='port',
                sample='oracle/mdo_oracle.c (x87 long double column loop of Reimer.py:476-686, OpenMP over rows) on one '
                       'F2 matrix n=%d, %.1f s, %d clamped columns; the Python reference itself measured 0.19-0.23 '
                       'GFLOP/s in the build container (BASELINE.md)' % (n, best, int(res['clamped'].sum())))
    return base, best, (res if keep_result else None)


def parity_against_oracle(ref, n=CPU_SAMPLE_N):
    """The GPU path through the public Python call on the matrix of the CPU sample, against the oracle's result for
    the same matrix: permutation, clamped set (Reimer.py:68), d, L element-wise, B = L D L^T (both composed on the
    device from the respective factors)."""
    import matrix_decomposition_b200 as matrix
    A, kw = cpu_sample_matrix(n)
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    d_ref = ref['d'].astype(np.float64)
    d_gpu = np.asarray(dec.d, dtype=np.float64)
    out = dict(n=n, against='oracle/mdo_oracle.c (x87 long double restatement of Reimer.py:476-686, pinned on the '
                            "reference's golden vectors)", call='approximation.positive_semidefinite.decomposition(A, **kw), default blocking',
               p_equal=bool(np.array_equal(np.asarray(dec.p), ref['p'])),
               clamped_equal=bool(np.array_equal(dec.clamped, ref['clamped'])),
               n_clamped=int(ref['clamped'].sum()),
               max_rel_d=float(np.max(np.abs(d_gpu - d_ref) / np.abs(d_ref))),
               max_abs_omega=float(np.max(np.abs(np.asarray(dec.omega, dtype=np.float64) - ref['omega'].astype(np.float64)))))
    B_gpu = dec.composed_matrix
    L_gpu = dec.L
    out['max_abs_dL'] = float(np.max(np.abs(L_gpu - ref['L']) / np.maximum(1.0, np.abs(ref['L']))))
    del L_gpu
    B_ref = matrix.decompositions.LDL_Decomposition(ref['L'], d_ref, p=ref['p']).composed_matrix
    out['relB'] = float(np.linalg.norm(B_gpu - B_ref) / np.linalg.norm(B_ref))
    out['tolerances'] = dict(max_rel_d=1e-10, max_abs_dL=1e-10, relB=1e-12)
    out['ok'] = bool(out['p_equal'] and out['clamped_equal'] and out['max_rel_d'] <= 1e-10 and out['max_abs_dL'] <= 1e-10
                     and out['relB'] <= 1e-12)
    return out


def cpu_baseline_subprocess(save_oracle=''):
    """cpu_baseline() in a fresh interpreter with a clean OpenMP environment: under torchrun the ranks inherit
    OMP_NUM_THREADS=1 and the x87 row loops ran 15 x slower than on the same box outside torchrun."""
    env = dict(os.environ)
    for k in ('OMP_NUM_THREADS', 'OMP_PROC_BIND', 'OMP_PLACES', 'GOMP_CPU_AFFINITY', 'KMP_AFFINITY'):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), '--cpu-baseline-json']
    if save_oracle:
        cmd += ['--save-oracle', save_oracle]
    out = subprocess.run(cmd, env=env, check=True, stdout=subprocess.PIPE, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    if os.environ.get('OMP_NUM_THREADS') and not os.environ.get('MDB200_REF_CHILD'):
        # launched under torchrun (it exports OMP_NUM_THREADS=1): run the CPU arm in a fresh interpreter with a clean
        # OpenMP environment so that it really uses all host cores, and relay its line
        env = dict(os.environ)
        for k in ('OMP_NUM_THREADS', 'OMP_PROC_BIND', 'OMP_PLACES', 'GOMP_CPU_AFFINITY', 'KMP_AFFINITY', 'RANK', 'WORLD_SIZE',
                  'LOCAL_RANK', 'LOCAL_WORLD_SIZE'):
            env.pop(k, None)
        env['MDB200_REF_CHILD'] = '1'
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--gpus', str(args.gpus),
                              '--steps', str(args.steps), '--warmup', str(args.warmup)], env=env, check=True,
                             stdout=subprocess.PIPE, text=True).stdout
        _RESULT.append(out.strip().splitlines()[-1])
        return 0
    n = CPU_SAMPLE_N
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base, dt, _ = cpu_baseline(n)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n ** 3 / 3 / (ms * 1e-3) / 1e9
    base['value'] = value
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='weak' if args.gpus == 1 else 'strong',
                vs_baseline=None, dtype='f64 (x87 long double accumulators, like the reference)', data='synthetic',
                config=dict(workload='matrix.approximate (modified LDL^T), F2 shifted GOE s=1.05, bounded CPU sample n=%d '
                                     'of the n=%d GPU workload' % (n, 65536 if args.gpus == 1 else 131072)),
                cpu_baseline=base,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)
    return 0


HBM_PEAK_GBS = 6547.5          # MEASURED_PEAKS.json (driver-written) when present, else this pool's figure


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:  # noqa: BLE001
        return HBM_PEAK_GBS


def _event_ms(stream, fn, reps=3, before=None):
    """Best device time of fn() over `reps` runs, CUDA events on the launching stream."""
    import torch
    best = None
    for _ in range(reps):
        if before is not None:
            before()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def extra_config2(ctx, stream, n=16384):
    """BASELINE configs[1]: matrix.decompose (LDL) of a dense SPD matrix n = 16384, then solve with 1 and 1024
    right-hand sides (dense/calculate.py:108-109, decompositions.py:983-991)."""
    import workloads
    import matrix_decomposition_b200 as matrix
    from matrix_decomposition_b200 import _native
    lib = ctx.lib
    A = workloads.lowrank_plus_diag(n, 3)
    p = np.argsort(-A.diagonal(), kind='stable').astype(np.int64)
    hbm = hbm_peak_gbs()
    out = dict(n=n, workload='F3 low-rank + diagonal SPD, decreasing_diagonal_values')
    # device time of the factorization (MDB_RULE_EXACT) with the permuted matrix resident
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'mdb_factor_create')
    bounds, keep = _native.make_bounds(_native.RULE_EXACT, 0.0, None, 0.0, 0.0)
    ms = _event_ms(stream, lambda: ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'mdb_factor_run'), reps=3,
                   before=lambda: ctx.check(lib.mdb_factor_set_matrix(f, _native.ptr(A), n, _native.ptr(p), _native.HOST),
                                            'mdb_factor_set_matrix'))
    lib.mdb_factor_destroy(f)
    out['decompose_device_ms'] = ms
    out['decompose_device_tflops'] = n ** 3 / 3 / (ms * 1e-3) / 1e12
    out['decompose_frac_fp64_peak'] = out['decompose_device_tflops'] / FP64_DMMA_PEAK_TFLOPS
    # public API, pageable numpy in, factor stays resident
    ts = []
    for _ in range(2):
        t = time.perf_counter()
        dec = matrix.decompose(A, permutation='decreasing_diagonal_values', return_type='LDL')
        ts.append(time.perf_counter() - t)
    out['decompose_api_s'] = min(ts)
    out['decompose_api_tflops'] = n ** 3 / 3 / min(ts) / 1e12
    rng = np.random.default_rng(5)
    b1 = rng.standard_normal(n)
    x1 = dec.solve(b1)     # first solve prepares the diagonal blocks
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        x1 = dec.solve(b1)
        ts.append(time.perf_counter() - t)
    ctx.set_profiling(True)
    dec.solve(b1)
    ms1 = ctx.phase_ms()['diag']     # device time of the two sweeps (CUDA events on the launching stream)
    ctx.set_profiling(False)
    out['solve_1rhs'] = dict(device_ms=ms1, achieved=8.0 * n * n / (ms1 * 1e-3) / 1e9, peak=hbm, unit='GB/s',
                             frac=8.0 * n * n / (ms1 * 1e-3) / 1e9 / hbm, algorithmic_bytes=8.0 * n * n,
                             api_s=min(ts), residual=float(np.max(np.abs(A @ x1 - b1)) / np.max(np.abs(b1))))
    k = 1024
    Bk = rng.standard_normal((n, k))
    ts = []
    for _ in range(2):
        t = time.perf_counter()
        Xk = dec.solve(Bk)
        ts.append(time.perf_counter() - t)
    ctx.set_profiling(True)
    dec.solve(Bk)
    msk = ctx.phase_ms()['diag']
    ctx.set_profiling(False)
    tf = 2.0 * n * n * k / (msk * 1e-3) / 1e12
    out['solve_1024rhs'] = dict(device_ms=msk, achieved=tf, peak=FP64_DMMA_PEAK_TFLOPS, unit='TFLOP/s',
                                frac=tf / FP64_DMMA_PEAK_TFLOPS, algorithmic_flop=2.0 * n * n * k, api_s=min(ts),
                                residual=float(np.max(np.abs(A @ Xk[:, :4] - Bk[:, :4])) / np.max(np.abs(Bk[:, :4]))))
    del dec
    return out


def extra_permute(ctx, stream, n=65536):
    """dense/permute.py:26 on the device: B[i, j] = A[p[i], p[j]] for a random p, device buffers in and out."""
    from matrix_decomposition_b200 import _native
    lib = ctx.lib
    hbm = hbm_peak_gbs()
    dA, dB = ctx.malloc(n * n * 8), ctx.malloc(n * n * 8)
    ctx.check(lib.mdb_generate_f2(ctx.handle, n, 1, 10.0, ctypes.c_void_p(dA), n), 'mdb_generate_f2')
    res = {}
    for name, p in (('random', np.random.default_rng(0).permutation(n).astype(np.int64)),
                    ('identity', np.arange(n, dtype=np.int64))):
        ms = _event_ms(stream, lambda: ctx.check(lib.mdb_permute_sym_f64(ctx.handle, n, ctypes.c_void_p(dA), n, _native.ptr(p),
                                                                        ctypes.c_void_p(dB), n, _native.DEVICE),
                                                 'mdb_permute_sym_f64'), reps=3)
        gbs = 16.0 * n * n / (ms * 1e-3) / 1e9
        moved = (32.0 if n > 8192 else 16.0) * n * n / (ms * 1e-3) / 1e9
        res[name] = dict(device_ms=ms, achieved=gbs, peak=hbm, unit='GB/s', frac=gbs / hbm, moved_gbs=moved,
                         moved_frac=moved / hbm)
    ctx.free(dA)
    ctx.free(dB)
    return dict(n=n, algorithmic_bytes=16.0 * n * n,
                note='achieved / frac: algorithmic bytes (read + write of the full matrix once) per second; above 8192 rows '
                     'the permutation runs as two passes of "gather rows, transpose" (k_gather_rows_transpose) because a '
                     'random p leaves a one-pass kernel with 8-byte gathers or scattered writes on one side (measured: '
                     '61-169 ms at n = 65536, DESIGN.md 4), so 32 n^2 bytes cross HBM: moved_gbs / moved_frac.  The '
                     'event-timed region includes the H2D copy of p (8 n bytes) and its validation on the host', **res)


def extra_config5(ctx, stream, batch=512, n=512, seed=0):
    """BASELINE configs[4], one GPU's share: `batch` independent n x n matrix.approximate problems (family F2), batched
    on the grid's z dimension.  Device time of the resident batched factorization and the public Python call
    decomposition_batched(As, ...) with pageable numpy arrays in and out."""
    import workloads
    import matrix_decomposition_b200 as matrix
    from matrix_decomposition_b200 import _native
    lib = ctx.lib
    rng = np.random.default_rng(seed)
    mu = S_SHIFT * np.sqrt(2 * n)
    G = rng.standard_normal((batch, n, n))
    As = (G + G.transpose(0, 2, 1)) / 2 + mu * np.eye(n)[None]
    del G
    kw = workloads.f2_bounds(n, S_SHIFT)
    ps = np.argsort(-np.diagonal(As, axis1=1, axis2=2), axis=1).astype(np.int64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, batch, ctypes.byref(f)), 'mdb_factor_create')
    l0 = [0]

    def before():
        ctx.check(lib.mdb_factor_set_matrix(f, _native.ptr(As), n, _native.ptr(ps), _native.HOST), 'mdb_factor_set_matrix')
        l0[0] = ctx.launch_count()
    ms = _event_ms(stream, lambda: ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'mdb_factor_run'), reps=4, before=before)
    launches = ctx.launch_count() - l0[0]
    lib.mdb_factor_destroy(f)
    psd = matrix.approximation.positive_semidefinite
    ts = []
    for _ in range(2):
        t = time.perf_counter()
        res = psd.decomposition_batched(As, **kw)
        ts.append(time.perf_counter() - t)
    sec = min(ts)
    tf = batch * n ** 3 / 3 / (ms * 1e-3) / 1e12
    return dict(batch=batch, n=n, device_ms=ms, device_matrices_per_s=batch / (ms * 1e-3), device_tflops=tf,
                frac_fp64_peak=tf / FP64_DMMA_PEAK_TFLOPS, launches=int(launches), e2e_s=sec,
                e2e_matrices_per_s=batch / sec, e2e_call='approximation.positive_semidefinite.decomposition_batched(As, **kw), '
                'pageable numpy in (%.1f GB) and out (L, d, omega, delta, clamped)' % (As.nbytes / 1e9),
                n_clamped_mean=float(res.clamped.sum(axis=1).mean()))


def extra_dynamic(n=16384):
    """Dynamic pivoting (permutation='maximal_stability', what the reference does for permutation=None, Reimer.py:415-416
    and :497-523) through the public Python call with a pageable numpy matrix: one k_dyn_column launch per column plus
    one K = 128 tensor-core update per panel.  Wall time, best of two after a warm-up call."""
    import workloads
    import matrix_decomposition_b200 as matrix
    A, kw = workloads.shifted_goe(n, 5, S_SHIFT)
    kw = dict(kw, permutation='maximal_stability')
    psd = matrix.approximation.positive_semidefinite
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        dec = psd.decomposition(A, **kw)
        ts.append(time.perf_counter() - t)
        ncl = int(dec.clamped.sum())
        del dec
    sec = min(ts[1:])
    return dict(n=n, method='maximal_stability', api_s=sec, us_per_column=sec / n * 1e6, tflops=n ** 3 / 3 / sec / 1e12,
                n_clamped=ncl, call='approximation.positive_semidefinite.decomposition(A, permutation="maximal_stability", '
                '**bounds), pageable numpy in, factor resident')


def run_single(args):
    import torch
    import matrix_decomposition_b200 as matrix  # noqa: F401  (fails loudly without libmdb200.so)
    from matrix_decomposition_b200 import _native

    n = args.n or 65536
    torch.cuda.set_device(0)
    ctx = _native.context(0)
    lib = ctx.lib
    stream = torch.cuda.Stream(priority=-1)   # high priority: panels overtake the bulk update on the side stream
    ctx.set_stream(stream.cuda_stream)
    mu = S_SHIFT * math.sqrt(2 * n)
    kw = f2_bounds(n)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(np.float64).eps), None, kw['max_diag_B'])
    seed = 20261019
    diag = np.empty(n)
    ctx.check(lib.mdb_generate_diag_f2(ctx.handle, n, seed, mu, _native.ptr(diag)), 'mdb_generate_diag_f2')
    p = np.argsort(-diag).astype(np.int64)   # 'decreasing_diagonal_values' (permute.py:46-65)

    # ---- HBM-resident leg: the unpermuted input A (34 GB at n = 65536) sits in HBM; one step = symmetric
    #      permutation gather into the working matrix + factorization + finalisation of L ----
    dA = ctx.malloc(n * n * 8)
    ctx.check(lib.mdb_generate_f2(ctx.handle, n, seed, mu, ctypes.c_void_p(dA), n), 'mdb_generate_f2')
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'mdb_factor_create')

    def step():
        ctx.check(lib.mdb_factor_set_matrix(f, ctypes.c_void_p(dA), n, _native.ptr(p), _native.DEVICE), 'mdb_factor_set_matrix')
        ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'mdb_factor_run')
        ctx.check(lib.mdb_factor_get(f, _native.TYPE_LDL, None, 0, _native.DEVICE, None, None, None, None, None),
                  'mdb_factor_get')   # finalises L in place (omega scaling, thresholding, unit diagonal); nothing is copied

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    times = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        with torch.cuda.stream(stream):
            e0.record(stream)
            step()
            e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        launches_per_step = ctx.launch_count() - l0
    clocks = sampler.stop()
    ms = float(np.mean(times))
    value = n ** 3 / 3 / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (k_gemm_tn_tma: the bulk trailing update, one launch per outer block),
    #      serialised profiling pass with CUDA events around every phase on the launching stream; the same pass
    #      leaves an un-finalised factor for the oracle-free identity of Reimer.py:947-957 at full size ----
    ctx.check(lib.mdb_factor_set_matrix(f, ctypes.c_void_p(dA), n, _native.ptr(p), _native.DEVICE), 'mdb_factor_set_matrix')
    ctx.free(dA)
    ctx.set_profiling(True)
    ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'mdb_factor_run')
    ctx.synchronize()
    ctx.set_profiling(False)
    phases = ctx.phase_ms()
    st = ctx.update_stats()
    res = ctypes.c_double()
    ctx.check(lib.mdb_factor_identity_residual(f, 2, 7, ctypes.byref(res)), 'mdb_factor_identity_residual')
    achieved = st['bulk_flop'] / (phases['trailing'] * 1e-3) / 1e12 if phases['trailing'] > 0 else None
    roofline = dict(bound='tensor', kernel='k_gemm_tn_tma (TMA-fed FP64 DMMA tile kernel; bulk trailing update, strict lower '
                                           'triangle, K = 128 x panels of the outer block)',
                    achieved=achieved, peak=FP64_DMMA_PEAK_TFLOPS, unit='TFLOP/s',
                    frac=(achieved / FP64_DMMA_PEAK_TFLOPS if achieved else None),
                    peak_source='measured FP64 DMMA peak (tools/fp64_peak.cu, profiles/r01_fp64_peak.jsonl); '
                                'MEASURED_PEAKS.json has no FP64 entry',
                    launches=st['bulk_launches'], avg_launch_ms=phases['trailing'] / max(st['bulk_launches'], 1),
                    algorithmic_flop_per_launch='K * rows * (rows - 1), rows = n - end of the outer block, K = its width '
                                                '(1024 / 512 / 256 / 128 while >= 32768 / 8192 / 4096 / 0 rows remain)',
                    algorithmic_flop_total=st['bulk_flop'], share_of_n3_over_3=st['bulk_flop'] / (n ** 3 / 3),
                    inner_update=dict(flop=st['inner_flop'], launches=st['inner_launches'], ms=phases['inner'],
                                      tflops=(st['inner_flop'] / (phases['inner'] * 1e-3) / 1e12 if phases['inner'] > 0 else None)),
                    phases_ms=phases, traffic=NCU_TRAFFIC, traffic_note=NCU_TRAFFIC_NOTE)
    lib.mdb_factor_destroy(f)

    # ---- end-to-end legs: C-ABI one-shot call with pinned host buffers; public Python call with pageable numpy ----
    e2e = None
    e2e_numpy = None
    if not args.no_e2e:
        e2e, e2e_numpy = run_e2e(ctx, n, seed, mu, p, bounds, args)
    ctx.trim()

    # ---- CPU baseline (the oracle port on a bounded sample) and, on the same matrix, parity of the GPU path ----
    base, _, ref = cpu_baseline(keep_result=True)
    parity = parity_against_oracle(ref)
    del ref
    extra = {}
    if not args.no_extra:
        for name, fn in (('config2_decompose_solve', lambda: extra_config2(ctx, stream)),
                         ('permute_symmetric', lambda: extra_permute(ctx, stream, min(n, 65536))),
                         ('config5_batched_512x512', lambda: extra_config5(ctx, stream)),
                         ('dynamic_pivoting', lambda: extra_dynamic())):
            try:
                extra[name] = fn()
            except Exception as e:  # noqa: BLE001  (an extra must not take the headline line down)
                extra[name] = dict(error='%s: %s' % (type(e).__name__, e))
            ctx.trim()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload='matrix.approximate (modified LDL^T, Reimer) n=%d float64, F2 shifted GOE s=1.05, '
                                     'min_diag_D=min_abs_value_D=0.015mu, max_diag_D=0.9mu, max_diag_B=2mu, '
                                     'decreasing_diagonal_values' % n,
                            n=n, l2='inputs (%.1f GB) far larger than L2' % (8e-9 * n * n),
                            timed_region='symmetric permutation gather of the HBM-resident A into the working matrix + '
                                         'factorization + in-place finalisation of L, CUDA events on the launching stream',
                            frac_of_fp64_dmma_peak=value / 1e3 / FP64_DMMA_PEAK_TFLOPS,
                            identity_residual=res.value, lookahead=os.environ.get('MDB200_LOOKAHEAD', '1'),
                            blocking='panels of 128 columns; outer blocks of 8/4/2/1 panels by remaining size'),
                clocks=clocks, gpu_launches=int(launches_per_step * args.steps), roofline=roofline, cpu_baseline=base,
                parity=parity, extra=extra)
    if e2e is not None:
        line['e2e'] = e2e
    if e2e_numpy is not None:
        line['e2e_numpy'] = e2e_numpy
    emit(line)
    return 0


def run_e2e(ctx, n, seed, mu, p, bounds, args):
    """mdb_approximate_f64 on pinned host buffers: H2D of A, factorization, D2H of L and the vectors."""
    from matrix_decomposition_b200 import _native
    lib = ctx.lib
    A = ctx.pinned_array((n, n))
    L = ctx.pinned_array((n, n))
    d = np.empty(n)
    omega = np.empty(n)
    delta = np.empty(n)
    clamped = np.zeros(n, dtype=np.uint8)
    # fill the host matrix with the same family (generated on the device, copied out once, untimed)
    dA = ctx.malloc(n * n * 8)
    ctx.check(lib.mdb_generate_f2(ctx.handle, n, seed, mu, ctypes.c_void_p(dA), n), 'mdb_generate_f2')
    ctx.check(lib.mdb_memcpy(ctx.handle, _native.ptr(A), ctypes.c_void_p(dA), n * n * 8, _native.HOST, _native.DEVICE),
              'mdb_memcpy')
    ctx.free(dA)
    times = []
    steps = max(1, min(args.steps, 2))
    for i in range(1 + steps):   # one warm-up call
        t = time.perf_counter()
        rc = lib.mdb_approximate_f64(ctx.handle, n, _native.ptr(A), n, _native.ptr(p), ctypes.byref(bounds),
                                     _native.TYPE_LDL, _native.ptr(L), n, _native.ptr(d), _native.ptr(omega),
                                     _native.ptr(delta), _native.ptr(clamped), _native.HOST, None)
        ctx.check(rc, 'mdb_approximate_f64')
        dt = time.perf_counter() - t
        if i > 0:
            times.append(dt)
    ok = bool(np.all(np.isfinite(d)) and L[n - 1, n - 1] == 1.0 and np.all(np.isfinite(L[n - 1])))
    _native.free_pinned(L)
    sec = float(np.mean(times))
    # bytes that actually cross PCIe: A is symmetric, mdb_factor_set_matrix uploads row blocks [r0, r0 + 2048) x
    # columns [r0, n) only (mdb_api.cu) and mirrors them on the device; plus the permutation vector
    rb = 2048
    h2d = (sum(min(rb, n - r0) * (n - r0) for r0 in range(0, n, rb)) if n >= 4096 else n * n) * 8 + n * 8
    e2e = dict(value=n ** 3 / 3 / sec / 1e9, unit=UNIT, h2d_bytes_per_step=int(h2d),
               h2d_note='upper block-triangle of the symmetric input (host array is n x n = {} bytes)'.format(n * n * 8),
               d2h_bytes_per_step=int(n * n * 8 + 3 * n * 8 + n), seconds_per_step=sec, steps=steps,
               call='mdb_approximate_f64 (include/mdb200.h) with pinned host A and L', results_finite=ok,
               n_clamped=int(clamped.sum()))
    # ---- the public Python call on a PAGEABLE numpy array (what north_star's drop-in user does) ----
    e2e_numpy = None
    need = 1.15 * n * n * 8
    if os.environ.get('BENCH_SKIP_NUMPY'):
        e2e_numpy = dict(skipped='BENCH_SKIP_NUMPY set')
    elif _mem_available_bytes() > need + (8 << 30):
        import matrix_decomposition_b200 as matrix
        An = np.empty((n, n))
        np.copyto(An, A)
        _native.free_pinned(A)
        A = None
        kw = f2_bounds(n)
        ts = []
        for i in range(2):     # one warm-up call
            t = time.perf_counter()
            dec = matrix.approximation.positive_semidefinite.decomposition(An, **kw)
            ts.append(time.perf_counter() - t)
            ncl = int(dec.clamped.sum())
            if i == 0:
                del dec
        e2e_numpy = dict(value=n ** 3 / 3 / ts[-1] / 1e9, unit=UNIT, seconds_per_step=ts[-1],
                         call='approximation.positive_semidefinite.decomposition(A, **kw) with a pageable numpy A; returns '
                              'the LDL_Decomposition with d / omega / delta / clamped on the host and L resident in HBM '
                              '(dec.solve / multiply use it there; dec.L downloads it on first access)',
                         h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(3 * n * 8 + n), n_clamped=ncl,
                         includes='numpy argsort of the diagonal, pageable -> pinned staging by the library\'s copy threads')
        del dec, An
    else:
        e2e_numpy = dict(skipped='host memory: %.0f GB available, %.0f GB needed for the pageable copy'
                                 % (_mem_available_bytes() / 1e9, need / 1e9))
    if A is not None:
        _native.free_pinned(A)
    return e2e, e2e_numpy


def _mem_available_bytes():
    try:
        for line in open('/proc/meminfo'):
            if line.startswith('MemAvailable:'):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def parity_distributed(ref_path, rank, world, device):
    """The multi-GPU path on the matrix of the CPU sample (n = 6144): vectors against the oracle on every rank's
    replicated copies, then the factor gathered to rank 0 and compared (a) with the oracle (L, B) and (b) BITWISE with
    the single-GPU factorization of the same matrix; plus the distributed solve against the single-GPU solve."""
    import matrix_decomposition_b200 as matrix
    from matrix_decomposition_b200 import distributed as mdist
    A, kw = cpu_sample_matrix()
    n = A.shape[0]
    fac = mdist.decomposition_distributed(A, **kw)
    rng = np.random.default_rng(3)
    b = rng.standard_normal((n, 2))
    x_dist = fac.solve(b)
    g = fac.gather(0)
    out = None
    if rank == 0:
        ref = dict(np.load(ref_path))
        d_ref = ref['d'].astype(np.float64)
        pos_omega = ref['omega'].astype(np.float64)[ref['p']]      # original index -> position order
        out = dict(n=n, world=world, call='distributed.decomposition_distributed(A, **kw) -> DistributedFactor; gather(0), solve(b)',
                   p_equal=bool(np.array_equal(fac.p, ref['p'])),
                   clamped_equal=bool(np.array_equal(fac.clamped, ref['clamped'])),
                   n_clamped=int(ref['clamped'].sum()),
                   max_rel_d=float(np.max(np.abs(fac.d - d_ref) / np.abs(d_ref))),
                   max_abs_omega=float(np.max(np.abs(fac.omega - pos_omega))))
        L = g.L
        out['max_abs_dL'] = float(np.max(np.abs(L - ref['L']) / np.maximum(1.0, np.abs(ref['L']))))
        B = g.composed_matrix
        B_ref = matrix.decompositions.LDL_Decomposition(ref['L'], d_ref, p=ref['p']).composed_matrix
        out['relB'] = float(np.linalg.norm(B - B_ref) / np.linalg.norm(B_ref))
        del B, B_ref
        one = matrix.approximation.positive_semidefinite.decomposition(A, **kw)     # single-GPU path on GPU 0
        out['bitwise_equal_single_gpu'] = bool(np.array_equal(L, one.L) and np.array_equal(fac.d, np.asarray(one.d, dtype=np.float64))
                                               and np.array_equal(fac.clamped, one.clamped)
                                               and np.array_equal(np.asarray(g.omega, dtype=np.float64), np.asarray(one.omega, dtype=np.float64))
                                               and np.array_equal(np.asarray(g.delta, dtype=np.float64), np.asarray(one.delta, dtype=np.float64)))
        x_one = one.solve(b)
        out['dist_solve_max_rel_err_vs_single_gpu'] = float(np.max(np.abs(x_dist - x_one)) / np.max(np.abs(x_one)))
        out['tolerances'] = dict(max_rel_d=1e-10, max_abs_dL=1e-10, relB=1e-12, dist_solve=1e-10)
        out['ok'] = bool(out['p_equal'] and out['clamped_equal'] and out['max_rel_d'] <= 1e-10 and out['max_abs_dL'] <= 1e-10
                         and out['relB'] <= 1e-12 and out['bitwise_equal_single_gpu']
                         and out['dist_solve_max_rel_err_vs_single_gpu'] <= 1e-10)
        del one, g, L
    fac.close()
    return out


def run_distributed(args):
    import torch
    import torch.distributed as dist
    from matrix_decomposition_b200 import distributed as mdist, _native

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', str(rank)))
    torch.cuda.set_device(local)
    ref_path = os.path.join(tempfile.gettempdir(), 'mdb200_oracle_sample_%d.npz' % os.getppid())
    base = cpu_baseline_subprocess(ref_path) if rank == 0 else None   # before the ranks start spinning on collectives
    mdist.init_process_group(local)
    n = args.n or 131072
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    out = mdist.bench_f2(n, S_SHIFT, steps=args.steps, warmup=args.warmup, seed=20261019, e2e=not args.no_e2e,
                         peak_tflops=FP64_DMMA_PEAK_TFLOPS, peak_hbm_gbs=hbm_peak_gbs())
    clocks = sampler.stop() if rank == 0 else None
    _native.context(local).trim()
    parity = None
    try:
        parity = parity_distributed(ref_path, rank, world, local)
    except Exception as e:  # noqa: BLE001
        parity = dict(error='%s: %s' % (type(e).__name__, e))
    extra = {}
    if not args.no_extra:
        # config 5: every rank factors its share of the 4096 matrices -- independent units, no data-path collective
        ctx = _native.context(local)
        stream = torch.cuda.Stream(priority=-1)
        ctx.set_stream(stream.cuda_stream)
        share = 4096 // world
        try:
            c5 = extra_config5(ctx, stream, batch=share, seed=rank)
            t = torch.tensor([c5['device_ms'], c5['e2e_s']], dtype=torch.float64, device=torch.device('cuda', local))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)      # reporting only
            c5.update(total_matrices=share * world, device_ms_max=float(t[0]), e2e_s_max=float(t[1]),
                      device_matrices_per_s_total=share * world / (float(t[0]) * 1e-3),
                      e2e_matrices_per_s_total=share * world / float(t[1]),
                      device_tflops_total=share * world * 512 ** 3 / 3 / (float(t[0]) * 1e-3) / 1e12,
                      frac_fp64_peak_total=share * world * 512 ** 3 / 3 / (float(t[0]) * 1e-3) / 1e12 / (FP64_DMMA_PEAK_TFLOPS * world))
            extra['config5_batched_4096x512x512'] = c5
        except Exception as e:  # noqa: BLE001
            extra['config5_batched_4096x512x512'] = dict(error='%s: %s' % (type(e).__name__, e))
        ctx.set_stream(None)
    if rank == 0:
        ms = out['ms_per_step']
        value = n ** 3 / 3 / (ms * 1e-3) / 1e9
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms, higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64',
                    data='synthetic',
                    config=dict(workload='matrix.approximate n=%d float64 (%.0f GB), F2 shifted GOE s=1.05, 1-D block-cyclic '
                                         'columns (block 128) over %d B200, panel exchange: %s, lookahead' %
                                         (n, 8e-9 * n * n, world, out.get('panel_exchange', 'NCCL broadcast')), n=n,
                                l2='inputs far larger than L2; matrix regenerated between steps',
                                frac_of_fp64_dmma_peak=value / 1e3 / (FP64_DMMA_PEAK_TFLOPS * world),
                                identity=out.get('identity'), n_clamped=out.get('n_clamped'),
                                e2e_skipped=out.get('e2e_skipped'), dist_solve=out.get('dist_solve')),
                    clocks=clocks, gpu_launches=out.get('gpu_launches'), roofline=out.get('roofline'),
                    cpu_baseline=base, e2e=out.get('e2e'), parity=parity, extra=extra)
        emit(line)
        try:
            os.unlink(ref_path)
        except OSError:
            pass
    dist.barrier()
    dist.destroy_process_group()
    return 0


class _QuietStdout:
    """Everything libraries print to stdout while the benchmark runs (e.g. NCCL's version banner) goes to stderr, so
    that stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def emit(line):
    """The one JSON line, on the real stdout."""
    _RESULT.append(json.dumps(line))


_RESULT = []


def main():
    with _QuietStdout():
        rc = _main()
    for line in _RESULT:
        print(line, flush=True)
    return rc


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', dest='n', type=int, default=0, help='override the matrix size (debugging only)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the config 2 / permutation / config 5 measurements')
    ap.add_argument('--cpu-baseline-json', action='store_true', help=argparse.SUPPRESS)
    ap.add_argument('--save-oracle', default='', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_json:
        base, _, ref = cpu_baseline(keep_result=bool(args.save_oracle))
        if args.save_oracle:
            np.savez(args.save_oracle, **{k: (v.astype(np.float64) if v.dtype == np.longdouble else v) for k, v in ref.items()})
        _RESULT.append(json.dumps(base))
        return 0
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 or world > 1:
        return run_distributed(args)
    return run_single(args)


if __name__ == '__main__':
    sys.exit(main())
