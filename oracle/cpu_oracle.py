"""ctypes front-end of the CPU oracle (oracle/mdo_oracle.c).

TEST INFRASTRUCTURE ONLY.  The product package (matrix_decomposition_b200/) never imports this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
and there only as the checker or the CPU baseline.  Parity is PINNED against the unmodified Python
reference through tests/golden/ (see tests/test_oracle.py).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD_DIR = os.path.join(_HERE, '_build')
_LIB_PATH = os.path.join(_BUILD_DIR, 'libmdo_oracle.so')
_SRC = os.path.join(_HERE, 'mdo_oracle.c')

PERM_STATIC = 0
PERM_MINIMAL_DIFFERENCE = 1
PERM_MAXIMAL_STABILITY = 2

_lib = None


def build(force=False):
    """Compile the C restatement with gcc (no -ffast-math: x87 long double semantics matter)."""
    os.makedirs(_BUILD_DIR, exist_ok=True)
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC)):
        return _LIB_PATH
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-fPIC', '-shared', '-std=gnu11', '-o', _LIB_PATH, _SRC, '-lm']
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, dbl, ldb, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_longdouble, ctypes.c_void_p
        L.mdo_cubic_real_roots_p.argtypes = [vp, vp]
        L.mdo_cubic_real_roots_p.restype = ctypes.c_int
        L.mdo_minimal_change_p.argtypes = [vp, vp]
        L.mdo_minimal_change_p.restype = ctypes.c_int
        L.mdo_approximate.argtypes = [i64, vp, i64, vp, ctypes.c_int, dbl, dbl, vp, ctypes.c_int, dbl, vp, vp, i64,
                                      vp, vp, vp, vp, vp, vp]
        L.mdo_approximate.restype = ctypes.c_int
        L.mdo_approximate_f64.argtypes = [i64, vp, i64, vp, dbl, dbl, dbl, dbl, vp, vp, i64, vp, vp, vp, vp, vp]
        L.mdo_approximate_f64.restype = ctypes.c_int
        L.mdo_psd_matrix.argtypes = [i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, ctypes.c_int, ctypes.c_int, vp]
        L.mdo_psd_matrix.restype = ctypes.c_int
        L.mdo_permute_sym.argtypes = [i64, vp, i64, vp, vp]
        L.mdo_permute_sym.restype = ctypes.c_int
        L.mdo_potrf_lower.argtypes = [i64, vp, i64, vp]
        L.mdo_potrf_lower.restype = i64
        L.mdo_solve.argtypes = [i64, i64, vp, vp, vp, vp, vp]
        L.mdo_solve.restype = ctypes.c_int
        L.mdo_multiply.argtypes = [i64, i64, vp, vp, vp, vp, vp]
        L.mdo_multiply.restype = ctypes.c_int
        _lib = L
    return _lib


def set_threads(n):
    """OpenMP threads of the row loops (timing harness only); returns the number now in use."""
    L = lib()
    L.mdo_set_threads.argtypes = [ctypes.c_int]
    L.mdo_set_threads.restype = ctypes.c_int
    return int(L.mdo_set_threads(int(n)))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# *** defaults of the reference (Reimer.py:396-408) *** #

def default_min_abs_value_D():
    """sqrt(eps(float128)) evaluated in long double, Reimer.py:398-402."""
    eps = np.finfo(np.longdouble).eps
    return max(eps ** 0.5, eps)


def default_min_abs_value_L():
    """eps(float64), Reimer.py:404-408."""
    return float(np.finfo(np.float64).eps)


# *** permutation vector (permute.py:10-87): pure numpy, the same calls as the reference *** #

def permutation_vector(A, method):
    n = A.shape[0]
    if method in (None, 'none'):
        return np.arange(n, dtype=np.int64)
    d = np.asarray(A).diagonal()
    if method in ('increasing_absolute_diagonal_values', 'decreasing_absolute_diagonal_values'):
        d = np.abs(d)
    if method in ('decreasing_diagonal_values', 'decreasing_absolute_diagonal_values'):
        d = -d
    elif method not in ('increasing_diagonal_values', 'increasing_absolute_diagonal_values'):
        raise ValueError(method)
    return np.argsort(d).astype(np.int64)


def cubic_real_roots(p, q):
    r = np.zeros(3, dtype=np.longdouble)
    pq = np.array([p, q], dtype=np.longdouble)
    k = lib().mdo_cubic_real_roots_p(_ptr(pq), _ptr(r))
    return [r[i] for i in range(k)]


def minimal_change(alpha, beta, gamma, min_diag_D, max_diag_D=math.inf, min_diag_B=-math.inf,
                   max_diag_B=math.inf, min_abs_value_D=None):
    """Scalar rule; returns (d, omega, f, clamped)."""
    if min_abs_value_D is None:
        min_abs_value_D = default_min_abs_value_D()
    out = np.zeros(3, dtype=np.longdouble)
    inp = np.array([alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, min_abs_value_D],
                   dtype=np.longdouble)
    c = lib().mdo_minimal_change_p(_ptr(inp), _ptr(out))
    return out[0], out[1], out[2], bool(c)


def _bounds_vectors(n, min_diag_B, max_diag_B):
    mB = np.asarray(-math.inf if min_diag_B is None else min_diag_B, dtype=np.float64)
    MB = np.asarray(math.inf if max_diag_B is None else max_diag_B, dtype=np.float64)
    if mB.ndim != MB.ndim:
        mB = np.broadcast_to(mB, (n,)).copy()
        MB = np.broadcast_to(MB, (n,)).copy()
    sB = 0 if mB.ndim == 0 else 1
    return np.ascontiguousarray(mB.reshape(-1)), np.ascontiguousarray(MB.reshape(-1)), sB


def approximate(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                min_abs_value_D=None, min_abs_value_L=None, permutation=None):
    """Restatement of Reimer._decomposition (dense, real).  Returns a dict with
    L (float64), d/omega/delta (longdouble), p (int64), clamped (bool, by position), f (longdouble)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    if permutation is None:
        permutation = 'maximal_stability'
    if isinstance(permutation, str):
        m = permutation.lower()
        if m == 'minimal_difference':
            mode, p = PERM_MINIMAL_DIFFERENCE, np.arange(n, dtype=np.int64)
        elif m == 'maximal_stability':
            mode, p = PERM_MAXIMAL_STABILITY, np.arange(n, dtype=np.int64)
        else:
            mode, p = PERM_STATIC, permutation_vector(A, m)
    else:
        mode, p = PERM_STATIC, np.array(permutation, dtype=np.int64)
    eps_ld = np.finfo(np.longdouble).eps
    if min_abs_value_D is None:
        mavD = default_min_abs_value_D()
    else:
        mavD = max(np.longdouble(float(min_abs_value_D)), eps_ld)
    mavL = default_min_abs_value_L() if min_abs_value_L is None else max(float(min_abs_value_L),
                                                                         default_min_abs_value_L())
    mB, MB, sB = _bounds_vectors(n, min_diag_B, max_diag_B)
    L = np.empty((n, n), dtype=np.float64)
    d = np.zeros(n, dtype=np.longdouble)
    omega = np.zeros(n, dtype=np.longdouble)
    delta = np.zeros(n, dtype=np.longdouble)
    f = np.zeros(n, dtype=np.longdouble)
    clamped = np.zeros(n, dtype=np.uint8)
    lib().mdo_approximate(n, _ptr(A), n, _ptr(p), mode,
                          math.nan if min_diag_D is None else float(min_diag_D),
                          math.inf if max_diag_D is None else float(max_diag_D),
                          _ptr(np.array([mavD], dtype=np.longdouble)), int(min_abs_value_D is not None and float(min_abs_value_D) >= eps_ld), mavL, _ptr(mB), _ptr(MB), sB,
                          _ptr(L), _ptr(d), _ptr(omega), _ptr(delta), _ptr(clamped), _ptr(f))
    return dict(L=L, d=d, p=p, omega=omega, delta=delta, clamped=clamped.astype(bool), f=f)


def approximate_f64(A, p, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                    min_abs_value_D=None, min_abs_value_L=None):
    """FP64 'fair twin' of the reference loop (static permutation only)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    p = np.ascontiguousarray(p, dtype=np.int64)
    mavD = float(default_min_abs_value_D()) if min_abs_value_D is None else float(min_abs_value_D)
    mavL = default_min_abs_value_L() if min_abs_value_L is None else float(min_abs_value_L)
    mB, MB, sB = _bounds_vectors(n, min_diag_B, max_diag_B)
    L = np.empty((n, n), dtype=np.float64)
    d = np.zeros(n)
    omega = np.zeros(n)
    delta = np.zeros(n)
    clamped = np.zeros(n, dtype=np.uint8)
    lib().mdo_approximate_f64(n, _ptr(A), n, _ptr(p), math.nan if min_diag_D is None else float(min_diag_D),
                              math.inf if max_diag_D is None else float(max_diag_D), mavD, mavL,
                              _ptr(mB), _ptr(MB), sB, _ptr(L), _ptr(d), _ptr(omega), _ptr(delta), _ptr(clamped))
    return dict(L=L, d=d, p=p, omega=omega, delta=delta, clamped=clamped.astype(bool))


def psd_matrix(A, res, min_diag_B=None, max_diag_B=None):
    """B of Reimer.positive_semidefinite_matrix from an approximate() result."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    mB, MB, sB = _bounds_vectors(n, min_diag_B, max_diag_B)
    B = np.empty((n, n))
    lib().mdo_psd_matrix(n, _ptr(A), n, _ptr(res['p']), _ptr(res['L']),
                         _ptr(np.ascontiguousarray(res['d'], dtype=np.longdouble)),
                         _ptr(np.ascontiguousarray(res['omega'], dtype=np.longdouble)),
                         _ptr(np.ascontiguousarray(res['delta'], dtype=np.longdouble)),
                         _ptr(mB), _ptr(MB), sB, int(min_diag_B is not None), int(max_diag_B is not None), _ptr(B))
    return B


def permute_symmetric(A, p):
    A = np.ascontiguousarray(A, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.int64)
    n = A.shape[0]
    B = np.empty((n, n))
    lib().mdo_permute_sym(n, _ptr(A), n, _ptr(p), _ptr(B))
    return B


def potrf_lower(A):
    """Returns (L, info) like scipy's potrf(lower=True, clean=True)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    L = np.empty((n, n))
    info = lib().mdo_potrf_lower(n, _ptr(A), n, _ptr(L))
    return L, int(info)


def decompose(A, permutation=None, return_type='LL'):
    """dense/calculate.py:15-126 + conversions decompositions.py:1248-1284.  Returns dict(L, d|None, p, info)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    if permutation is None or isinstance(permutation, str):
        p = permutation_vector(A, permutation)
    else:
        p = np.array(permutation, dtype=np.int64)
    Ap = permute_symmetric(A, p)
    L, info = potrf_lower(Ap)
    if info != 0 or return_type == 'LL':
        return dict(L=L, d=None, p=p, info=info)
    dg = L.diagonal().copy()
    Lu = L / dg[None, :]
    np.fill_diagonal(Lu, 1.0)
    return dict(L=Lu, d=dg * dg, p=p, info=info)


def solve(L, d, p, b):
    """decompositions.py:983-991 (LDL, d given) / :1345-1352 (LL, d None)."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    n = L.shape[0]
    p = np.ascontiguousarray(np.arange(n) if p is None else p, dtype=np.int64)
    b = np.asarray(b, dtype=np.float64)
    b2 = np.ascontiguousarray(b.reshape(n, -1))
    x = np.empty_like(b2)
    dd = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
    lib().mdo_solve(n, b2.shape[1], _ptr(L), _ptr(dd), _ptr(p), _ptr(b2), _ptr(x))
    return x.reshape(b.shape)


def multiply(L, d, p, b):
    L = np.ascontiguousarray(L, dtype=np.float64)
    n = L.shape[0]
    p = np.ascontiguousarray(np.arange(n) if p is None else p, dtype=np.int64)
    b = np.asarray(b, dtype=np.float64)
    b2 = np.ascontiguousarray(b.reshape(n, -1))
    x = np.empty_like(b2)
    dd = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
    lib().mdo_multiply(n, b2.shape[1], _ptr(L), _ptr(dd), _ptr(p), _ptr(b2), _ptr(x))
    return x.reshape(b.shape)
