#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native matrix-decomposition hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size SIZE]

A "step" is one modified LDL^T factorization (matrix.approximate /
approximation.positive_semidefinite.decomposition) of one synthetic dense symmetric float64 matrix of
family F2 (GOE + mu I, mu = 1.05 sqrt(2n), bounds min_diag_D = min_abs_value_D = 0.015 mu,
max_diag_D = 0.9 mu, max_diag_B = 2 mu, permutation 'decreasing_diagonal_values').

  N = 1 : n = 65536 (34 GB), BASELINE.json configs[2].  `value` = n^3/3 flop / device time of the
          factorization with the matrix already resident in HBM; `e2e` = the same metric through the
          C-ABI one-shot call mdb_approximate_f64 with pinned HOST buffers (H2D of A, D2H of L, d, omega,
          delta, clamped inside the timed region).
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
# (profiles/r01_ncu_summary.txt); None until captured for the current kernel
NCU_TRAFFIC = 46.01e9   # bulk update part (b), grid 234240 tiles, 109.3 ms: read 30.94 GB + write 15.08 GB (profiles/r01b_ncu_summary.txt)
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


def cpu_baseline(n=CPU_SAMPLE_N, reps=1):
    """The oracle port of the reference's column loop, all host cores (OpenMP), bounded sample."""
    import workloads
    from oracle import cpu_oracle as orc
    A, kw = workloads.shifted_goe(n, 4242, S_SHIFT)
    # all host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    cores = orc.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    orc.approximate(A[:256, :256].copy(), **f2_bounds(256))  # build + warm the library
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        res = orc.approximate(A, **kw)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return dict(value=n ** 3 / 3 / best / 1e9, unit=UNIT, cores=cores, kind='port',
                sample='oracle/mdo_oracle.c (x87 long double column loop of Reimer.py:476-686, OpenMP over rows) on one '
                       'F2 matrix n=%d, %.1f s, %d clamped columns; the Python reference itself measured 0.19-0.23 '
                       'GFLOP/s in the build container (BASELINE.md)' % (n, best, int(res['clamped'].sum()))), best


def cpu_baseline_subprocess():
    """cpu_baseline() in a fresh interpreter with a clean OpenMP environment: under torchrun the ranks inherit
    OMP_NUM_THREADS=1 and the x87 row loops ran 15 x slower than on the same box outside torchrun."""
    env = dict(os.environ)
    for k in ('OMP_NUM_THREADS', 'OMP_PROC_BIND', 'OMP_PLACES', 'GOMP_CPU_AFFINITY', 'KMP_AFFINITY'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--cpu-baseline-json'], env=env, check=True,
                         stdout=subprocess.PIPE, text=True).stdout
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
        base, dt = cpu_baseline(n)
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

    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'mdb_factor_create')

    def regenerate():
        ctx.check(lib.mdb_factor_generate_f2(f, seed, mu, _native.ptr(p)), 'mdb_factor_generate_f2')

    def factor():
        ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'mdb_factor_run')

    # ---- HBM-resident leg: W (34 GB at n = 65536) is regenerated outside the timed region ----
    for _ in range(args.warmup):
        regenerate()
        factor()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ctx.launch_count()
    times = []
    for _ in range(args.steps):
        regenerate()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        with torch.cuda.stream(stream):
            e0.record(stream)
            factor()
            e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        launches_per_step = ctx.launch_count() - l0
    clocks = sampler.stop()
    ms = float(np.mean(times))
    value = n ** 3 / 3 / (ms * 1e-3) / 1e9

    # ---- parity guard at full size: the oracle-free identity of Reimer.py:947-957 on the device ----
    res = ctypes.c_double()
    ctx.check(lib.mdb_factor_identity_residual(f, 2, 7, ctypes.byref(res)), 'mdb_factor_identity_residual')

    # ---- roofline of the dominant kernel (k_gemm_tn_tma: the bulk trailing update, one launch per outer block),
    #      serialised profiling pass with CUDA events around every phase on the launching stream ----
    regenerate()
    ctx.set_profiling(True)
    factor()
    ctx.synchronize()
    ctx.set_profiling(False)
    phases = ctx.phase_ms()
    st = ctx.update_stats()
    achieved = st['bulk_flop'] / (phases['trailing'] * 1e-3) / 1e12 if phases['trailing'] > 0 else None
    roofline = dict(bound='tensor', kernel='k_gemm_tn_tma (TMA-fed FP64 DMMA tile kernel; bulk trailing update, strict lower '
                                           'triangle, K = 128 x panels of the outer block)',
                    achieved=achieved, peak=FP64_DMMA_PEAK_TFLOPS, unit='TFLOP/s',
                    frac=(achieved / FP64_DMMA_PEAK_TFLOPS if achieved else None),
                    peak_source='measured FP64 DMMA peak (tools/fp64_peak.cu, profiles/r01_fp64_peak.jsonl); '
                                'MEASURED_PEAKS.json has no FP64 entry',
                    launches=st['bulk_launches'], avg_launch_ms=phases['trailing'] / max(st['bulk_launches'], 1),
                    algorithmic_flop_per_launch='K * rows * (rows - 1), rows = n - end of the outer block, K = its width '
                                                '(1024 / 512 / 256 / 128 while >= 32768 / 16384 / 8192 / 0 rows remain)',
                    algorithmic_flop_total=st['bulk_flop'], share_of_n3_over_3=st['bulk_flop'] / (n ** 3 / 3),
                    inner_update=dict(flop=st['inner_flop'], launches=st['inner_launches'], ms=phases['inner'],
                                      tflops=(st['inner_flop'] / (phases['inner'] * 1e-3) / 1e12 if phases['inner'] > 0 else None)),
                    phases_ms=phases, traffic=NCU_TRAFFIC,
                    traffic_note='ncu --set full capture of the LARGEST bulk launch (part (b) of the first outer block, grid '
                                 '234240 tiles, 109.3 ms, DMMA pipe 95.4 % active): dram read 30.94 GB + write 15.08 GB vs '
                                 'algorithmic 31.7 GB (C triangle read + written 30.7 GB, operands 1.0 GB); the excess are '
                                 'operand re-reads that missed L2 (profiles/r01b_ncu_summary.txt)')
    lib.mdb_factor_destroy(f)

    # ---- end-to-end leg: C-ABI one-shot call with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(ctx, n, seed, mu, p, bounds, args)

    base, _ = cpu_baseline()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload='matrix.approximate (modified LDL^T, Reimer) n=%d float64, F2 shifted GOE s=1.05, '
                                     'min_diag_D=min_abs_value_D=0.015mu, max_diag_D=0.9mu, max_diag_B=2mu, '
                                     'decreasing_diagonal_values' % n,
                            n=n, l2='inputs (%.1f GB) far larger than L2; matrix regenerated between steps' % (8e-9 * n * n),
                            frac_of_fp64_dmma_peak=value / 1e3 / FP64_DMMA_PEAK_TFLOPS,
                            identity_residual=res.value, lookahead=os.environ.get('MDB200_LOOKAHEAD', '1'),
                            blocking='panels of 128 columns; outer blocks of 8/4/2/1 panels by remaining size'),
                clocks=clocks, gpu_launches=int(launches_per_step * args.steps), roofline=roofline, cpu_baseline=base)
    if e2e is not None:
        line['e2e'] = e2e
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
    _native.free_pinned(A)
    _native.free_pinned(L)
    sec = float(np.mean(times))
    # bytes that actually cross PCIe: A is symmetric, mdb_factor_set_matrix uploads row blocks [r0, r0 + 2048) x
    # columns [r0, n) only (mdb_api.cu) and mirrors them on the device; plus the permutation vector
    rb = 2048
    h2d = (sum(min(rb, n - r0) * (n - r0) for r0 in range(0, n, rb)) if n >= 4096 else n * n) * 8 + n * 8
    return dict(value=n ** 3 / 3 / sec / 1e9, unit=UNIT, h2d_bytes_per_step=int(h2d),
                h2d_note='upper block-triangle of the symmetric input (host array is n x n = {} bytes)'.format(n * n * 8),
                d2h_bytes_per_step=int(n * n * 8 + 3 * n * 8 + n), seconds_per_step=sec, steps=steps,
                call='mdb_approximate_f64 (include/mdb200.h) with pinned host A and L', results_finite=ok,
                n_clamped=int(clamped.sum()))


def run_distributed(args):
    import torch
    import torch.distributed as dist
    from matrix_decomposition_b200 import distributed as mdist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', str(rank)))
    torch.cuda.set_device(local)
    base = cpu_baseline_subprocess() if rank == 0 else None   # before the ranks start spinning on collectives
    mdist.init_process_group(local)
    n = args.n or 131072
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    out = mdist.bench_f2(n, S_SHIFT, steps=args.steps, warmup=args.warmup, seed=20261019, e2e=not args.no_e2e,
                         peak_tflops=FP64_DMMA_PEAK_TFLOPS)
    if rank == 0:
        out['clocks'] = sampler.stop()
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
                                e2e_skipped=out.get('e2e_skipped')),
                    clocks=out.get('clocks'), gpu_launches=out.get('gpu_launches'), roofline=out.get('roofline'),
                    cpu_baseline=base, e2e=out.get('e2e'))
        emit(line)
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
    ap.add_argument('--cpu-baseline-json', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_json:
        _RESULT.append(json.dumps(cpu_baseline()[0]))
        return 0
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 or world > 1:
        return run_distributed(args)
    return run_single(args)


if __name__ == '__main__':
    sys.exit(main())
