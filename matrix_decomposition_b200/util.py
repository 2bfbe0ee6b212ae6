"""Array coercion helpers with the reference's copy semantics (matrix/util.py:10-36, :66-80, :146-165)."""
import numpy as np

from . import errors


def as_matrix_or_array(A, check_ndim_values=None, dtype=None, copy=True):
    A = np.array(A, dtype=dtype, copy=True if copy else None)
    if check_ndim_values is not None and A.ndim not in check_ndim_values:
        raise ValueError('The number of dimensions of the input should be in '
                         '{} but it is {}.'.format(check_ndim_values, A.ndim))
    return A


def as_matrix_or_vector(A, dtype=None, copy=True):
    return as_matrix_or_array(A, check_ndim_values=(1, 2), dtype=dtype, copy=copy)


def as_matrix(A, dtype=None, copy=True):
    return as_matrix_or_array(A, check_ndim_values=(2,), dtype=dtype, copy=copy)


def as_vector(A, dtype=None, copy=True):
    A = np.array(A, dtype=dtype, copy=True if copy else None)
    if A.ndim != 1:
        raise ValueError('The number of dimensions of the input should be 1 but it is {}.'.format(A.ndim))
    return A


def is_equal(A, B):
    return A.shape == B.shape and not np.any(A != B)


def is_almost_equal(A, B, rtol=1e-05, atol=1e-08):
    return A.shape == B.shape and np.allclose(A, B, rtol=rtol, atol=atol)


def check_square_matrix(A):
    if len(A.shape) != 2 or A.shape[0] != A.shape[1]:
        raise errors.MatrixNotSquareError(A)


def is_finite(A):
    return bool(np.all(np.isfinite(A)))


def check_finite(A, check_finite=True):
    if check_finite and not is_finite(A):
        raise errors.MatrixNotFiniteError(A)


def conjugate_transpose(x, copy=False):
    return x.conj().transpose()


def block_diag(A, B):
    n = A.shape[0]
    m = B.shape[0]
    return np.block([[A, np.zeros((n, m))], [np.zeros((m, n)), B]])
