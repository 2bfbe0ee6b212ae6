"""Modified LDL^T approximation (Reimer's algorithm) behind the reference's API.

Reference: matrix/approximation/positive_semidefinite/Reimer.py (`R.py` below).  This module is the
host side only: argument validation and defaults exactly as R.py:273-457, the static permutation
vector (numpy argsort of the diagonal, like the reference), and the call into libmdb200.so, where the
column loop R.py:476-686 runs as a blocked right-looking factorization on the B200.
"""
import ctypes
import math

import numpy as np

from ... import _native, constants, decompositions, errors, permute
from ... import logger

MINIMAL_DIFFERENCE_PERMUTATION_METHOD = 'minimal_difference'
""" Permutation method supported by the decomposition and the positive_semidefinite_matrix algorithm. """
MAXIMAL_STABILITY_PERMUTATION_METHOD = 'maximal_stability'
""" Permutation method supported by the decomposition and the positive_semidefinite_matrix algorithm. """
APPROXIMATION_ONLY_PERMUTATION_METHODS = (MINIMAL_DIFFERENCE_PERMUTATION_METHOD, MAXIMAL_STABILITY_PERMUTATION_METHOD)
""" Permutation methods supported only by the decomposition and the positive_semidefinite_matrix algorithm. """

# The reference keeps d / omega / delta in numpy float128 (R.py:396,460-464).  The device computes in
# float64; results are returned as np.longdouble arrays so dtype-dependent behaviour downstream
# (e.g. the singularity threshold finfo(d.dtype).resolution, decompositions.py:949-951) is unchanged.
RESULT_DTYPE = np.longdouble


def _check_float_scalar_or_vector(f, n, f_name, default_value, minus_inf_okay=False, plus_inf_okay=False):
    """R.py:291-322"""
    if f is not None:
        try:
            f = np.asarray(f)
        except TypeError as original_error:
            error = ValueError(f'{f_name} must a scalar value or a vector but it is {f}.')
            logger.error(error)
            raise error from original_error
        if not (f.ndim == 0 or (f.ndim == 1 and f.shape[0] == n)):
            error = ValueError(f'{f_name} must be a scalar or a vector with the same '
                               f'dimension as A but its shape is {f.shape}.')
            logger.error(error)
            raise error
        if not (np.all(np.isreal(f))):
            error = ValueError(f'{f_name} must be real valued but it is {f}.')
            logger.error(error)
            raise error
        ok = np.isfinite(f)
        if plus_inf_okay:
            ok = np.logical_or(ok, f == math.inf)
        if minus_inf_okay:
            ok = np.logical_or(ok, f == -math.inf)
        if not np.all(ok):
            error = ValueError(f'{f_name} must be finite'
                               + (' or minus infinity' if minus_inf_okay else '')
                               + (' or plus infinity' if plus_inf_okay else '')
                               + f' but it is {f}.')
            logger.error(error)
            raise error
    else:
        f = np.asarray(default_value)
    return f


def _check_float_scalar(f, f_name, default_value=None, lower_bound=None, upper_bound=None,
                        minus_inf_okay=False, plus_inf_okay=False):
    """R.py:324-354"""
    if f is not None:
        try:
            f = float(f)
        except TypeError as original_error:
            error = ValueError(f'{f_name} must a float value but it is {f}.')
            logger.error(error)
            raise error from original_error
        if not (math.isfinite(f) or (minus_inf_okay and f == -math.inf) or (plus_inf_okay and f == math.inf)):
            error = ValueError(f'{f_name} must be finite'
                               + (' or minus infinity' if minus_inf_okay else '')
                               + (' or plus infinity' if plus_inf_okay else '')
                               + f' but it is {f}.')
            logger.error(error)
            raise error
        if lower_bound is not None and f < lower_bound:
            error = ValueError(f'{f_name} must be greater or equal to {lower_bound} but it is {f}.')
            logger.error(error)
            raise error
        if upper_bound is not None and f > upper_bound:
            error = ValueError(f'{f_name} must be less or equal to {upper_bound} but it is {f}.')
            logger.error(error)
            raise error
    else:
        f = default_value
    return f


def _prepare(A, min_diag_B, max_diag_B, min_diag_D, max_diag_D, min_abs_value_D, min_abs_value_L, permutation):
    """Validation and defaults of R.py:273-457 for the dense real case."""
    A = np.asarray(A)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise errors.MatrixNotSquareError(A)
    n = A.shape[0]
    if np.issubdtype(A.dtype, np.complexfloating):
        gamma = A.diagonal()
        if not np.all(np.isreal(gamma)):
            raise errors.MatrixComplexDiagonalValueError(A, i=np.where(~np.isreal(gamma))[0][0])
        raise ValueError('complex matrices are outside the dense real (float64) path of this package.')

    min_diag_B = _check_float_scalar_or_vector(min_diag_B, n, 'min_diag_B', -math.inf, minus_inf_okay=True)
    max_diag_B = _check_float_scalar_or_vector(max_diag_B, n, 'max_diag_B', math.inf, plus_inf_okay=True)
    min_diag_D = _check_float_scalar(min_diag_D, 'min_diag_D', default_value=None, lower_bound=0)
    max_diag_D = _check_float_scalar(max_diag_D, 'max_diag_D', default_value=math.inf, plus_inf_okay=True)
    if not (max(np.max(min_diag_B), min_diag_D if min_diag_D is not None else 0)
            <= min(np.min(max_diag_B), max_diag_D)):  # R.py:368-373
        error = ValueError(('The values of min_diag_B and min_diag_D must be lower or equal '
                            'to the values of max_diag_B and max_diag_D.'))
        logger.error(error)
        raise error

    d_eps = np.finfo(np.longdouble).eps  # R.py:396-408
    min_abs_value_D = _check_float_scalar(min_abs_value_D, 'min_abs_value_D', default_value=d_eps**0.5, lower_bound=0)
    min_abs_value_D = max(min_abs_value_D, d_eps)
    L_eps = np.finfo(np.float64).eps
    min_abs_value_L = _check_float_scalar(min_abs_value_L, 'min_abs_value_L', default_value=L_eps,
                                          lower_bound=0, upper_bound=1)
    min_abs_value_L = max(min_abs_value_L, L_eps)

    dynamic = 0
    if permutation is None:  # R.py:415-416
        permutation = MAXIMAL_STABILITY_PERMUTATION_METHOD
    if isinstance(permutation, str):
        permutation_method = permutation.lower()
        supported = constants.UNIVERSAL_PERMUTATION_METHODS + APPROXIMATION_ONLY_PERMUTATION_METHODS
        if permutation_method not in supported:
            error = ValueError(f'Permutation method {permutation_method} is unknown. Only the '
                               f'following methods are supported {supported}.')
            logger.error(error)
            raise error
        if permutation_method in APPROXIMATION_ONLY_PERMUTATION_METHODS:
            if (permutation_method == MINIMAL_DIFFERENCE_PERMUTATION_METHOD
                    and min_diag_D is not None and min_diag_D <= 0):  # R.py:437-441
                raise ValueError(f'The permutation method {MINIMAL_DIFFERENCE_PERMUTATION_METHOD} is '
                                 f'only available if min_diag_D is greater zero.')
            # dynamic pivoting starts from the identity (R.py:443) and is resolved on the device
            p = None
            dynamic = 1 if permutation_method == MINIMAL_DIFFERENCE_PERMUTATION_METHOD else 2
        else:
            p = permute.permutation_vector(A, permutation_method=permutation_method)
    else:
        p = np.asanyarray(permutation)
        if p.ndim != 1 or p.shape[0] != n:
            error = ValueError(f'Permutation vector must have same length as the dimensions of '
                               f'A. Its shape is {p.shape} and the shape of A is {A.shape}.')
            logger.error(error)
            raise error
    return (A, n, p, dynamic, min_diag_B, max_diag_B, min_diag_D, max_diag_D, float(min_abs_value_D),
            float(min_abs_value_L))


def _type_code(out_type_str):
    return _native.TYPE_NONE if out_type_str is None else _native.TYPE_CODES[out_type_str]


def _run(A, out_type_str, keep_factor, download_L=True, **kw):
    """download_L=False leaves the n x n factor in HBM (returned Lout is None; needs keep_factor)."""
    A, n, p, dynamic, min_diag_B, max_diag_B, min_diag_D, max_diag_D, mavD, mavL = _prepare(A, **kw)
    A64 = np.ascontiguousarray(A, dtype=np.float64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, min_diag_D, max_diag_D, mavD, mavL, min_diag_B, max_diag_B)
    Lout = np.empty((n, n), dtype=np.float64) if download_L else None
    d = np.empty(n, dtype=np.float64)
    omega = np.empty(n, dtype=np.float64)
    delta = np.empty(n, dtype=np.float64)
    clamped = np.zeros(n, dtype=np.uint8)
    handle = ctypes.c_void_p()
    ctx = _native.context()
    P = _native.ptr
    if dynamic:
        p = np.empty(n, dtype=np.int64)
        rc = ctx.lib.mdb_approximate_dynamic_f64(ctx.handle, n, P(A64), n, dynamic, ctypes.byref(bounds),
                                                 _type_code(out_type_str), P(Lout), n, P(p), P(d), P(omega),
                                                 P(delta), P(clamped), _native.HOST,
                                                 ctypes.byref(handle) if keep_factor else None)
        ctx.check(rc, 'mdb_approximate_dynamic_f64')
        p = p.astype(np.min_scalar_type(n)) if n > 0 else p   # dtype of the reference's arange (R.py:443)
    else:
        p64 = np.ascontiguousarray(p, dtype=np.int64)
        rc = ctx.lib.mdb_approximate_f64(ctx.handle, n, P(A64), n, P(p64), ctypes.byref(bounds),
                                         _type_code(out_type_str), P(Lout), n, P(d), P(omega), P(delta),
                                         P(clamped), _native.HOST, ctypes.byref(handle) if keep_factor else None)
        ctx.check(rc, 'mdb_approximate_f64')
    del keep
    factor = _native.Factor(ctx, handle) if keep_factor else None
    return Lout, d, p, omega, delta, clamped.astype(bool), factor


def decomposition(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                  min_abs_value_D=None, min_abs_value_L=None, permutation=None, overwrite_A=False,
                  return_type=None):
    """
    Computes an approximative decomposition of a matrix with the specified properties (R.py:712-816).

    Same parameters, defaults, exceptions and return value as the reference: a decomposition object of
    type `return_type` (default 'LDL') carrying the extra attributes `.omega` and `.delta`
    (R.py:810-812; indexed by ORIGINAL index, float128 like the reference's).  Additionally `.clamped`
    holds the boolean mask (by position) of the columns whose (d, omega) left the unclamped branch of
    R.py:68.  `A` is never modified (`overwrite_A` only documents permission).

    The dynamic pivoting methods 'maximal_stability' (the default for permutation=None, R.py:415-416) and
    'minimal_difference' run as an unblocked right-looking factorization with a device-wide arg-min per
    column (mdb_approximate_dynamic_f64); they need n^2/2 rule evaluations and are meant for n up to ~10^4.
    """
    supported_return_types = constants.DECOMPOSITION_TYPES
    if return_type is not None and return_type not in supported_return_types:
        error = ValueError(f'Unkown return type {return_type}. Only values in '
                           f'{supported_return_types} are supported.')
        logger.error(error)
        raise error
    type_str = constants.LDL_DECOMPOSITION_TYPE if return_type is None else return_type
    # the factor stays in HBM: dec.L / dec.LD download it on first access, solve / multiply never need to
    _, d64, p, omega, delta, clamped, factor = _run(
        A, type_str, True, download_L=False, min_diag_B=min_diag_B, max_diag_B=max_diag_B, min_diag_D=min_diag_D,
        max_diag_D=max_diag_D, min_abs_value_D=min_abs_value_D, min_abs_value_L=min_abs_value_L,
        permutation=permutation)
    d = d64.astype(RESULT_DTYPE)
    n = len(d)
    if type_str == constants.LL_DECOMPOSITION_TYPE:
        for i, d_i in enumerate(d):  # decompositions.py:903-908 (cannot happen: min_diag_D >= 0)
            if d_i < 0:
                raise errors.NoDecompositionPossibleWithProblematicSubdecompositionError(
                    decompositions.LDL_Decomposition(np.eye(len(d)), d, p), constants.LL_DECOMPOSITION_TYPE, p[i])
        dec = decompositions.LL_Decomposition(None, p=p)
        dec._attach_lazy_matrix(factor, n, np.sqrt(d64))
    elif type_str == constants.LDL_DECOMPOSITION_COMPRESSED_TYPE:
        dec = decompositions.LDL_DecompositionCompressed(None, p=p)
        dec._attach_lazy_matrix(factor, n, d64)
    else:
        dec = decompositions.LDL_Decomposition(None, d, p=p)
        dec._attach_lazy_matrix(factor, n, np.ones(n))
    dec.omega = omega.astype(RESULT_DTYPE)
    dec.delta = delta.astype(RESULT_DTYPE)
    dec.clamped = clamped
    return dec


def positive_semidefinite_matrix(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                                 min_abs_value_D=None, min_abs_value_L=None, permutation=None, overwrite_A=False):
    """
    Computes an approximation `B` of `A` which has a LDL^H decomposition with the specified properties
    (R.py:819-992): B[i, i] = A[i, i] + delta[i] clipped to [min_diag_B, max_diag_B] (R.py:924-945),
    B[i, j] = A[i, j] * omega[later of (i, j) in elimination order] (R.py:947-957), with the inner-product
    fallback where d == 0 (R.py:958-975).

    Both the factorization and the formation of B (an n^2 Python double loop in the reference) run on the
    device: the factor is left un-finalised in HBM (its upper triangle still holds the permuted A) and
    mdb_factor_psd_matrix writes B in original index order.
    """
    _, d, p, omega, delta, clamped, factor = _run(
        A, None, True, download_L=False, min_diag_B=min_diag_B, max_diag_B=max_diag_B,
        min_diag_D=min_diag_D, max_diag_D=max_diag_D, min_abs_value_D=min_abs_value_D,
        min_abs_value_L=min_abs_value_L, permutation=permutation)
    n = len(d)
    B = np.empty((n, n), dtype=np.float64)
    ctx = factor.ctx

    def bound(x):
        if x is None:
            return None, 0
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 0:   # (np.ascontiguousarray would silently turn a scalar into a 1-vector)
            return np.full(1, float(x)), 0
        return np.ascontiguousarray(x), 1
    mB, s1 = bound(min_diag_B)
    MB, s2 = bound(max_diag_B)
    stride = max(s1, s2)
    if stride:   # mixed scalar / vector bounds: broadcast the scalar one
        if mB is not None and s1 == 0:
            mB = np.full(n, mB[0])
        if MB is not None and s2 == 0:
            MB = np.full(n, MB[0])
    rc = ctx.lib.mdb_factor_psd_matrix(factor.handle, _native.ptr(mB), _native.ptr(MB), stride, _native.ptr(B), n, _native.HOST)
    ctx.check(rc, 'mdb_factor_psd_matrix')
    factor.close()
    return B
