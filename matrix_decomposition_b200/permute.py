"""Permutation vectors and symmetric permutation (reference: matrix/permute.py, matrix/dense/permute.py)."""
import numpy as np

from . import _native, constants, errors


def permutation_vector(A, permutation_method=None):
    """Permutation vector for a permutation method (permute.py:10-87, dense methods).

    O(n log n) host work on the diagonal only, done with the very numpy calls of the reference
    (np.arange with min_scalar_type, np.argsort of +-diag / +-|diag|), so the vector -- including the
    order among equal diagonal entries -- is identical to the reference's on the same numpy.
    """
    DIAGONAL_VALUES_PERMUATION_METHODS = constants.DIAGONAL_VALUES_PERMUTATION_METHODS
    A = np.asarray(A) if not hasattr(A, 'diagonal') else A
    if permutation_method == constants.NO_PERMUTATION_METHOD:
        n = A.shape[0]
        p = np.arange(n, dtype=np.min_scalar_type(n))
    elif permutation_method in DIAGONAL_VALUES_PERMUATION_METHODS:
        d = A.diagonal()
        if permutation_method in (constants.INCREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD,
                                  constants.DECREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD):
            d = np.abs(d)
        if permutation_method in (constants.INCREASING_DIAGONAL_VALUES_PERMUTATION_METHOD,
                                  constants.DECREASING_DIAGONAL_VALUES_PERMUTATION_METHOD):
            if np.iscomplexobj(d):
                if np.all(np.isreal(d)):
                    d = d.real
                else:
                    raise errors.MatrixComplexDiagonalValueError(A)
        if permutation_method in (constants.DECREASING_DIAGONAL_VALUES_PERMUTATION_METHOD,
                                  constants.DECREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD):
            d = -d
        p = np.argsort(d)
    else:
        raise ValueError(('Permutation method {} is unknown. Only the following methods are supported '
                          'for dense matrices: {}.').format(permutation_method, constants.UNIVERSAL_PERMUTATION_METHODS))
    return p


def invert_permutation_vector(p):
    """permute.py:90-93"""
    p_inverse = np.empty_like(p)
    p_inverse[p] = np.arange(len(p))
    return p_inverse


def concatenate_permutation_vectors(p_previous, p_next):
    """permute.py:96-102"""
    if p_previous is None:
        return p_next
    elif p_next is None:
        return p_previous
    else:
        return p_previous[p_next]


def symmetric(A, p):
    """`B[i, j] == A[p[i], p[j]]` (permute.py:105-129, dense/permute.py:26): a coalesced gather kernel on the
    device.  Returns a new array of A's dtype kind (float64); `A` is never modified."""
    if p is None:
        return A
    A = np.asarray(A)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise errors.MatrixNotSquareError(A)
    if np.iscomplexobj(A):
        raise ValueError('complex matrices are outside the dense real (float64) path of this package.')
    n = A.shape[0]
    p64 = np.ascontiguousarray(p, dtype=np.int64)
    if p64.ndim != 1 or p64.shape[0] != n:
        raise ValueError('Permutation vector must have same length as the dimensions of A.')
    A64 = np.ascontiguousarray(A, dtype=np.float64)
    B = np.empty((n, n), dtype=np.float64)
    ctx = _native.context()
    rc = ctx.lib.mdb_permute_sym_f64(ctx.handle, n, _native.ptr(A64), n, _native.ptr(p64), _native.ptr(B), n, _native.HOST)
    ctx.check(rc, 'mdb_permute_sym_f64')
    return B
