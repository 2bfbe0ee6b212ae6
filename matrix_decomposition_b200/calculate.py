"""matrix.decompose (reference: matrix/calculate.py:11-75, matrix/dense/calculate.py:15-126)."""
import ctypes

import numpy as np

from . import _native, constants, decompositions, errors, permute, util


def _resolve_permutation(A, permutation, supported_methods):
    """dense/calculate.py:79-95: method name -> vector, or validate an explicit vector."""
    if permutation is None:
        permutation = constants.NO_PERMUTATION_METHOD
    if isinstance(permutation, str):
        permutation_method = permutation.lower()
        if permutation_method not in supported_methods:
            raise ValueError(('Permutation method {} is unknown. Only the following methods are supported {}.'
                              ).format(permutation_method, supported_methods))
        p = permute.permutation_vector(A, permutation_method)
    else:
        p = np.asanyarray(permutation)
        if p.ndim != 1 or p.shape[0] != A.shape[0]:
            raise ValueError(('Permutation vactor must have same length as the dimensions of A. '
                              'Its shape is {} and the shape of A is {}.').format(p.shape, A.shape))
    return p


_HOST_FINITE_SCAN_MAX = 1 << 20   # elements


def _make_decomposition(type_str, Lout, d, p):
    if type_str == constants.LL_DECOMPOSITION_TYPE:
        return decompositions.LL_Decomposition(Lout, p=p)
    if type_str == constants.LDL_DECOMPOSITION_COMPRESSED_TYPE:
        return decompositions.LDL_DecompositionCompressed(Lout, p=p)
    return decompositions.LDL_Decomposition(Lout, d, p=p)


def decompose(A, permutation=None, return_type=None, check_finite=True, overwrite_A=False):
    """
    Computes a decomposition of a dense positive definite matrix (calculate.py:11-75).

    Same arguments, return value and exceptions as the reference.  The permutation gather, the
    factorization (blocked right-looking LDL^T with an FP64 tensor-core trailing update, the GPU
    counterpart of LAPACK potrf at dense/calculate.py:108-109) and the conversion to `return_type`
    run on the device; `A` is never modified (so `overwrite_A` only documents permission).

    Raises
    ------
    matrix.errors.NoDecompositionPossibleWithProblematicSubdecompositionError
        If `A` is not positive definite; `problematic_leading_principal_submatrix_index` is LAPACK's info - 1
        (dense/calculate.py:115-119).
    matrix.errors.MatrixNotSquareError, matrix.errors.MatrixNotFiniteError
    """
    A = np.asanyarray(A)
    util.check_square_matrix(A)
    if np.iscomplexobj(A):
        raise ValueError('complex matrices are outside the dense real (float64) path of this package.')
    # dense/calculate.py:76.  Small inputs are scanned here; for large ones the scan (a pass over n^2 host doubles
    # plus an n^2 temporary) rides on the device-side gather instead (checked right after the call below).
    host_scan = A.size <= _HOST_FINITE_SCAN_MAX
    util.check_finite(A, check_finite=check_finite and host_scan)
    p = _resolve_permutation(A, permutation, constants.UNIVERSAL_PERMUTATION_METHODS)
    supported_return_type = constants.DECOMPOSITION_TYPES
    if return_type is not None and return_type not in supported_return_type:
        raise ValueError(('Unkown decomposition type {}. Only values in {} are supported.'
                          ).format(return_type, supported_return_type))
    type_str = constants.LL_DECOMPOSITION_TYPE if return_type is None else return_type  # potrf yields LL

    n = A.shape[0]
    A64 = np.ascontiguousarray(A, dtype=np.float64)
    p64 = np.ascontiguousarray(p, dtype=np.int64)
    d = np.empty(n, dtype=np.float64)
    info = ctypes.c_int64(0)
    handle = ctypes.c_void_p()
    ctx = _native.context()
    # The factor stays in HBM (L = NULL: no n x n download); .L / .LD fetch it on first access.
    rc = ctx.lib.mdb_decompose_f64(ctx.handle, n, _native.ptr(A64), n, _native.ptr(p64), _native.TYPE_CODES[type_str],
                                   None, n, _native.ptr(d), ctypes.byref(info), _native.HOST, ctypes.byref(handle))
    # dense/calculate.py:76: the finite check of A, evaluated on the device while A was gathered
    if check_finite and ctx.lib.mdb_ctx_last_input_nonfinite(ctx.handle):
        if handle:
            ctx.lib.mdb_factor_destroy(handle)
        raise errors.MatrixNotFiniteError(A)
    if rc == _native.MDB_ERR_NOT_POSITIVE_DEFINITE:
        bad_index = int(info.value) - 1
        # like dense/calculate.py:112-119: hand back the partial Cholesky factor with NaN at the bad pivot
        # (rare path: run again, this time downloading the partial factor)
        Lout = np.empty((n, n), dtype=np.float64)
        ctx.lib.mdb_decompose_f64(ctx.handle, n, _native.ptr(A64), n, _native.ptr(p64), _native.TYPE_CODES[type_str],
                                  _native.ptr(Lout), n, _native.ptr(d), ctypes.byref(info), _native.HOST, None)
        Lpart = np.tril(Lout)
        Lpart[bad_index, bad_index] = np.nan
        sub = decompositions.LL_Decomposition(Lpart, p=p)
        raise errors.NoDecompositionPossibleWithProblematicSubdecompositionError(
            A, constants.LL_DECOMPOSITION_TYPE, bad_index, sub)
    ctx.check(rc, 'mdb_decompose_f64')
    factor = _native.Factor(ctx, handle)
    if type_str == constants.LL_DECOMPOSITION_TYPE:
        dec = decompositions.LL_Decomposition(None, p=p)
        dec._attach_lazy_matrix(factor, n, np.sqrt(d))
    elif type_str == constants.LDL_DECOMPOSITION_COMPRESSED_TYPE:
        dec = decompositions.LDL_DecompositionCompressed(None, p=p)
        dec._attach_lazy_matrix(factor, n, d)
    else:
        dec = decompositions.LDL_Decomposition(None, d, p=p)
        dec._attach_lazy_matrix(factor, n, np.ones(n))
    return dec


def is_positive_semidefinite(A, check_finite=True):
    """calculate.py:78-120"""
    try:
        decomposition = decompose(A, permutation=constants.NO_PERMUTATION_METHOD, check_finite=check_finite)
    except (errors.NoDecompositionPossibleError, errors.MatrixComplexDiagonalValueError,
            errors.MatrixNotFiniteError, errors.MatrixNotSquareError):
        return False
    return decomposition.is_positive_semidefinite()


def is_positive_definite(A, check_finite=True):
    """calculate.py:123-165"""
    try:
        decomposition = decompose(A, permutation=constants.NO_PERMUTATION_METHOD, check_finite=check_finite)
    except (errors.NoDecompositionPossibleError, errors.MatrixComplexDiagonalValueError,
            errors.MatrixNotFiniteError, errors.MatrixNotSquareError):
        return False
    return decomposition.is_positive_definite()


def is_invertible(A, check_finite=True):
    """calculate.py:168-208"""
    try:
        decomposition = decompose(A, permutation=constants.NO_PERMUTATION_METHOD, check_finite=check_finite)
    except (errors.NoDecompositionPossibleError, errors.MatrixComplexDiagonalValueError,
            errors.MatrixNotFiniteError, errors.MatrixNotSquareError):
        return False
    return decomposition.is_invertible()


def solve(A, b, overwrite_b=False, check_finite=True):
    """calculate.py:211-262.  (The reference's own version raises TypeError, calculate.py:258 passes
    keywords that decomposition.solve does not accept; this one decomposes and solves.)"""
    util.check_finite(np.asarray(b), check_finite=check_finite)
    try:
        decomposition = decompose(A, check_finite=check_finite)
    except errors.NoDecompositionPossibleError:
        raise errors.MatrixSingularError(np.asarray(A))
    try:
        return decomposition.solve(b)
    except errors.DecompositionSingularError as e:
        raise errors.MatrixSingularError(np.asarray(A)) from e
