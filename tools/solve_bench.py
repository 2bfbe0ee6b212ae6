"""BASELINE config 2: matrix.decompose (LDL) of a dense SPD float64 matrix n = 16384 on one B200, then
decomposition.solve with 1 and 1024 right-hand sides, through the public Python API (host numpy arrays in and
out).  Prints one JSON object: factorization GFLOP/s (n^3/3), 1-rhs solve as GB/s of the algorithmic 8 n^2
bytes against the measured HBM peak, 1024-rhs solve as TFLOP/s (2 n^2 k) against the FP64 DMMA peak, and the
residuals.  --cpu also times scipy's potrf / two trtrs on the host cores (the reference's own calls,
dense/calculate.py:108-109, dense/util.py:24-29)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import workloads  # noqa: E402
import matrix_decomposition_b200 as matrix  # noqa: E402
from matrix_decomposition_b200 import _native  # noqa: E402

HBM_PEAK_GBS = 6547.5      # MEASURED_PEAKS.json (driver-written, this pool)
FP64_PEAK_TFLOPS = 37.1    # tools/fp64_peak.cu


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16384
    ctx = _native.context()
    rng = np.random.default_rng(5)
    A = workloads.lowrank_plus_diag(n, 3)
    out = dict(n=n)
    # decompose through the public API (H2D of A, D2H of L inside)
    t, dec = best_of(lambda: matrix.decompose(A, permutation='decreasing_diagonal_values', return_type='LDL'), 2)
    out['decompose_api_s'] = t
    out['decompose_api_gflops'] = n ** 3 / 3 / t / 1e9
    b1 = rng.standard_normal(n)
    x1 = dec.solve(b1)                       # first call uploads nothing (factor is resident) but prepares the blocks
    l0 = ctx.launch_count()
    t1, x1 = best_of(lambda: dec.solve(b1), 5)
    out['solve1_launches'] = (ctx.launch_count() - l0) // 5
    out['solve1_s'] = t1
    out['solve1_gbs'] = 8.0 * n * n / t1 / 1e9
    out['solve1_frac_hbm'] = out['solve1_gbs'] / HBM_PEAK_GBS
    ctx.set_profiling(True)
    dec.solve(b1)
    ms = ctx.phase_ms()['diag']      # device time of the two sweeps (CUDA events on the launching stream)
    out['solve1_device_ms'] = ms
    out['solve1_device_gbs'] = 8.0 * n * n / (ms * 1e-3) / 1e9
    out['solve1_device_frac_hbm'] = out['solve1_device_gbs'] / HBM_PEAK_GBS
    ctx.set_profiling(False)
    out['solve1_residual'] = float(np.max(np.abs(A @ x1 - b1)) / np.max(np.abs(b1)))
    k = 1024
    Bk = rng.standard_normal((n, k))
    tk, Xk = best_of(lambda: dec.solve(Bk), 3)
    out['solve1024_s'] = tk
    out['solve1024_tflops'] = 2.0 * n * n * k / tk / 1e12
    out['solve1024_frac_fp64_peak'] = out['solve1024_tflops'] / FP64_PEAK_TFLOPS
    ctx.set_profiling(True)
    dec.solve(Bk)
    ms = ctx.phase_ms()['diag']
    out['solve1024_device_ms'] = ms
    out['solve1024_device_tflops'] = 2.0 * n * n * k / (ms * 1e-3) / 1e12
    out['solve1024_device_frac_fp64_peak'] = out['solve1024_device_tflops'] / FP64_PEAK_TFLOPS
    ctx.set_profiling(False)
    out['solve1024_residual'] = float(np.max(np.abs(A @ Xk[:, :4] - Bk[:, :4])) / np.max(np.abs(Bk[:, :4])))
    t = time.perf_counter()
    L = dec.L                                   # first access: the factor comes home (pageable numpy array)
    out['L_download_s'] = time.perf_counter() - t
    out['L_download_gbs'] = 8.0 * n * n / out['L_download_s'] / 1e9
    del L
    if '--cpu' in sys.argv:
        import scipy.linalg
        p = np.asarray(dec.p)
        Ap = A[p[:, None], p[None, :]]
        t = time.perf_counter()
        Lc = scipy.linalg.cholesky(Ap, lower=True, check_finite=False)
        out['cpu_potrf_s'] = time.perf_counter() - t
        out['cpu_potrf_gflops'] = n ** 3 / 3 / out['cpu_potrf_s'] / 1e9

        def cpu_solve(B):
            y = scipy.linalg.solve_triangular(Lc, B[p], lower=True)
            z = scipy.linalg.solve_triangular(Lc.T, y, lower=False)
            return z
        out['cpu_solve1_s'], _ = best_of(lambda: cpu_solve(b1), 2)
        out['cpu_solve1024_s'], _ = best_of(lambda: cpu_solve(Bk), 1)
        out['cpu_cores'] = os.cpu_count()
    print(json.dumps(out))
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'solve_bench.json'), 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
