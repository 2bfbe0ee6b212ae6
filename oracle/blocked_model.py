"""numpy model of the blocked right-looking algorithm that the CUDA kernels implement.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  It exists so that the restatement the GPU
runs -- unscaled rows Lu, Schur complement S kept next to the original A, the mixing identity
c = (1 - w) A[:, i] + w S[:, i], the all-dropped shortcut, TRSM against the scaled + thresholded
diagonal block -- can be validated on CPU against the oracle (tests/test_blocked_model.py) before
any GPU time is spent.  The scalar rule is the very header the kernels compile (csrc/md_rule.h), built
for the host by tests/host_rule_shim.cpp.

Restatement, for a static permutation, Ap = A[p][:, p]:
  S   = Ap - Lu D Lu^T on the trailing block (lower triangle of W), Ap kept in the upper triangle
  per column K:  (d, w) = rule(alpha_K, beta_K, gamma_K)
                 r_j = Ap[j, K]                      if w == 0 or w * rowmax_K < min_abs_value_L
                     = (1 - w) Ap[j, K] + w S0[j, K]  otherwise          (S0: S at the start of the panel)
                 Lu[j, K] = (r_j - sum_{m in panel, m < K} Lu[j, m] d_m Lt[K, m]) / d
                 with Lt[K, m] = thresh(w Lu[K, m]) the scaled, thresholded row (Reimer.py:553-560)
                 alpha_j += Lu[j, K]^2 d,  beta_j += 2 Ap[j, K]^2,  rowmax_j = max(rowmax_j, |Lu[j, K]|)
  trailing:      S[j, j'] -= sum_{K in panel} Lu[j, K] d_K Lu[j', K]
  output:        L[j, m] = thresh(w_j Lu[j, m]), unit diagonal
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_SHIM_SRC = os.path.join(_REPO, 'tests', 'host_rule_shim.cpp')
_SHIM_SO = os.path.join(_HERE, '_build', 'host_rule.so')
_shim = None

RULE_REIMER = 0
RULE_EXACT = 1


def rule_shim():
    global _shim
    if _shim is None:
        os.makedirs(os.path.dirname(_SHIM_SO), exist_ok=True)
        hdr = os.path.join(_REPO, 'matrix_decomposition_b200', 'csrc', 'md_rule.h')
        if (not os.path.exists(_SHIM_SO)
                or os.path.getmtime(_SHIM_SO) < max(os.path.getmtime(_SHIM_SRC), os.path.getmtime(hdr))):
            subprocess.run(['g++', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-o', _SHIM_SO, _SHIM_SRC],
                           check=True)
        s = ctypes.CDLL(_SHIM_SO)
        dbl, vp = ctypes.c_double, ctypes.c_void_p
        s.mdh_cubic.argtypes = [dbl, dbl, vp]
        s.mdh_cubic.restype = ctypes.c_int
        s.mdh_minimal_change.argtypes = [dbl] * 8 + [vp]
        s.mdh_minimal_change.restype = ctypes.c_int
        s.mdh_minimal_change_split.argtypes = [dbl] * 8 + [vp]
        s.mdh_minimal_change_split.restype = ctypes.c_int
        s.mdh_minimal_change_full.argtypes = [dbl] * 8 + [vp]
        s.mdh_minimal_change_full.restype = ctypes.c_int
        s.mdh_upper_clamp_shortcut.argtypes = [dbl] * 8 + [vp]
        s.mdh_upper_clamp_shortcut.restype = ctypes.c_int
        s.mdh_upper_clamp_shortcut_cheap.argtypes = [dbl] * 8
        s.mdh_upper_clamp_shortcut_cheap.restype = ctypes.c_int
        s.mdh_upper_clamp_shortcut_row.argtypes = [dbl] * 8
        s.mdh_upper_clamp_shortcut_row.restype = ctypes.c_int
        s.mdh_decide.argtypes = [ctypes.c_int] + [dbl] * 8 + [vp]
        s.mdh_decide.restype = ctypes.c_int
        _shim = s
    return _shim


def decide(rule, min_diag_D, max_diag_D, mavD, alpha, beta, gamma, mB, MB):
    out = np.zeros(3)
    c = rule_shim().mdh_decide(rule, min_diag_D, max_diag_D, mavD, alpha, beta, gamma, mB, MB,
                               out.ctypes.data_as(ctypes.c_void_p))
    return out[0], out[1], out[2], bool(c)


def blocked_factor(A, p, rule=RULE_REIMER, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                   min_abs_value_D=None, min_abs_value_L=None, nb=128):
    """Returns dict(L, d, omega, delta, clamped, p, info) with d/clamped by position, omega/delta by
    original index (like the reference).  info = first failing pivot + 1 for RULE_EXACT, else 0."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    p = np.asarray(p, dtype=np.int64)
    mavD = float(np.finfo(np.longdouble).eps ** 0.5) if min_abs_value_D is None else max(float(min_abs_value_D),
                                                                                         float(np.finfo(np.longdouble).eps))
    eps64 = float(np.finfo(np.float64).eps)
    mavL = eps64 if min_abs_value_L is None else max(float(min_abs_value_L), eps64)
    mD = math.nan if min_diag_D is None else float(min_diag_D)
    MD = math.inf if max_diag_D is None else float(max_diag_D)
    mBv = np.broadcast_to(np.asarray(-math.inf if min_diag_B is None else min_diag_B, dtype=np.float64), (n,))[p]
    MBv = np.broadcast_to(np.asarray(math.inf if max_diag_B is None else max_diag_B, dtype=np.float64), (n,))[p]

    W = A[p[:, None], p[None, :]].copy()
    gamma = W.diagonal().copy()
    alpha = np.zeros(n)
    beta = np.zeros(n)
    rowmax = np.zeros(n)
    d = np.zeros(n)
    omega = np.zeros(n)
    clamped = np.zeros(n, dtype=bool)
    delta = np.zeros(n)
    info = 0

    for k0 in range(0, n, nb):
        kb = min(nb, n - k0)
        k1 = k0 + kb
        # right-hand sides for all rows j >= k0 (diagonal block rows and panel rows alike)
        Apan = W[k0:k1, k0:].T.copy()      # Apan[j - k0, k] = Ap[j, k0 + k]  (upper triangle; symmetric)
        Span = W[k0:, k0:k1].copy()        # Span[j - k0, k] = S0[j, k0 + k]  (valid for j > k0 + k)
        X = np.zeros((n - k0, kb))         # unscaled Lu[j, k0 + k]
        Lt = np.zeros((kb, kb))
        for k in range(kb):
            K = k0 + k
            dk, wk, _, cl = decide(rule, mD, MD, mavD, alpha[K], beta[K], gamma[K], mBv[K], MBv[K])
            d[K], omega[K], clamped[K] = dk, wk, cl
            delta[K] = (dk - gamma[K]) + (wk * wk * alpha[K] if wk != 0 else 0.0)  # Reimer.py:541-545
            if rule == RULE_EXACT and cl and info == 0:
                info = K + 1
            drop = (wk == 0) or (wk * rowmax[K] < mavL)
            if wk != 0:
                t = wk * X[k, :k]
                t[np.abs(t) < mavL] = 0
                Lt[k, :k] = t
            u = d[k0:K] * Lt[k, :k]
            a = Apan[k + 1:, k]
            if drop:
                r = a.copy()
            else:
                r = (1 - wk) * a + wk * Span[k + 1:, k]
            if k > 0:
                # sequential accumulation in m, like a device thread would do it
                acc = np.zeros(n - K - 1)
                for m in range(k):
                    acc += X[k + 1:, m] * u[m]
                r = r - acc
            x = r / dk if dk != 0 else np.zeros_like(r)
            X[k + 1:, k] = x
            alpha[K + 1:] += x * x * dk
            beta[K + 1:] += 2 * a * a
            rowmax[K + 1:] = np.maximum(rowmax[K + 1:], np.abs(x))
        # store Lu (strict lower part of the diagonal block + panel rows)
        for k in range(kb):
            W[k0 + k + 1:, k0 + k] = X[k + 1:, k]
        # trailing update (lower triangle only; the diagonal of S is never read)
        if k1 < n:
            Xp = X[kb:, :]
            Y = Xp * d[k0:k1][None, :]
            upd = Xp @ Y.T
            il = np.tril_indices(n - k1, -1)
            T = W[k1:, k1:]
            T[il] -= upd[il]
    # finalize
    L = np.tril(W, -1) * omega[:, None]
    L[np.abs(L) < mavL] = 0
    L[np.diag_indices(n)] = 1
    om_orig = np.zeros(n)
    om_orig[p] = omega
    de_orig = np.zeros(n)
    de_orig[p] = delta
    return dict(L=L, d=d, omega=om_orig, delta=de_orig, clamped=clamped, p=p, info=info, W=W, alpha=alpha, beta=beta)


def dynamic_delayed_factor(A, method, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                           min_abs_value_D=None, min_abs_value_L=None, nb=128):
    """numpy model of csrc/kernels_dynamic.cuh: dynamic pivoting ('maximal_stability' / 'minimal_difference',
    Reimer.py:497-523) with the Schur-complement updates delayed per panel of nb columns.

    Per column i: rule for every remaining row, arg-min by the reference's key, symmetric swap i <-> j* (rows of the
    pending operands X, Y included), column i read as  S[j, i] + sum_c X[j, c] Y[i, c]  over the panel's earlier
    columns, mixed with the original A by omega (the all-dropped shortcut as in the static path), divided by d;
    alpha / beta / rowmax updated for every row below.  After the panel one product updates the trailing S.
    Returns dict(L, d, omega, delta, clamped, p) with d / clamped by position, omega / delta by original index."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    mavD = float(np.finfo(np.longdouble).eps ** 0.5) if min_abs_value_D is None else max(float(min_abs_value_D),
                                                                                         float(np.finfo(np.longdouble).eps))
    eps64 = float(np.finfo(np.float64).eps)
    mavL = eps64 if min_abs_value_L is None else max(float(min_abs_value_L), eps64)
    mD = math.nan if min_diag_D is None else float(min_diag_D)
    MD = math.inf if max_diag_D is None else float(max_diag_D)
    mB = np.broadcast_to(np.asarray(-math.inf if min_diag_B is None else min_diag_B, dtype=np.float64), (n,)).copy()
    MB = np.broadcast_to(np.asarray(math.inf if max_diag_B is None else max_diag_B, dtype=np.float64), (n,)).copy()
    W = A.copy()                   # lower: S / Lu, upper: A (both follow the swaps)
    p = np.arange(n)
    gamma = A.diagonal().copy()
    alpha, beta, rowmax = np.zeros(n), np.zeros(n), np.zeros(n)
    d, omega, delta = np.zeros(n), np.zeros(n), np.zeros(n)
    clamped = np.zeros(n, dtype=bool)
    PX, PY = np.zeros((n, nb)), np.zeros((n, nb))

    def swap_vec(v, i, j):
        v[i], v[j] = v[j], v[i]

    for pb in range(0, n, nb):
        pe = min(pb + nb, n)
        for i in range(pb, pe):
            kc = i - pb
            best = None
            for j in range(i, n):
                dj, wj, fj, cj = decide(RULE_REIMER, mD, MD, mavD, alpha[j], beta[j], gamma[j], mB[j], MB[j])
                key = (fj, -dj, wj, j) if method == 'minimal_difference' else (-dj, fj, wj, j)
                if best is None or key < best[0]:
                    best = (key, dj, wj, cj, j)
            _, di, wi, ci, js = best
            if js != i:
                for v in (gamma, alpha, beta, rowmax, mB, MB, p):
                    swap_vec(v, i, js)
                PX[[i, js], :kc] = PX[[js, i], :kc]
                PY[[i, js], :kc] = PY[[js, i], :kc]
                for m in range(n):          # S(a, b) at W[max, min], A(a, b) at W[min, max]
                    if m == i or m == js:
                        continue
                    lo_i, lo_j = (max(i, m), min(i, m)), (max(js, m), min(js, m))
                    W[lo_i], W[lo_j] = W[lo_j], W[lo_i]
                    up_i, up_j = (min(i, m), max(i, m)), (min(js, m), max(js, m))
                    W[up_i], W[up_j] = W[up_j], W[up_i]
            d[i], omega[i], clamped[i] = di, wi, ci
            delta[i] = (di - gamma[i]) + (wi * wi * alpha[i] if wi != 0 else 0.0)
            drop = (wi == 0) or (wi * rowmax[i] < mavL)
            a = W[i, i + 1:].copy()
            r = a.copy()
            if not drop:
                s = W[i + 1:, i] + PX[i + 1:, :kc] @ PY[i, :kc]
                r = (1 - wi) * a + wi * s
            x = r / di if di != 0 else np.zeros_like(r)
            W[i + 1:, i] = x
            alpha[i + 1:] += x * x * di
            beta[i + 1:] += 2 * a * a
            rowmax[i + 1:] = np.maximum(rowmax[i + 1:], np.abs(x))
            PX[i + 1:, kc] = x
            PY[i + 1:, kc] = -x * di
        if pe < n:
            upd = PX[pe:, :pe - pb] @ PY[pe:, :pe - pb].T
            il = np.tril_indices(n - pe, -1)
            T = W[pe:, pe:]
            T[il] += upd[il]
    L = np.tril(W, -1) * omega[:, None]
    L[np.abs(L) < mavL] = 0
    L[np.diag_indices(n)] = 1
    om_orig, de_orig = np.zeros(n), np.zeros(n)
    om_orig[p] = omega
    de_orig[p] = delta
    return dict(L=L, d=d, omega=om_orig, delta=de_orig, clamped=clamped, p=p)
