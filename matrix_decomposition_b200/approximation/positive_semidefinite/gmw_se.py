"""GMW and SE families of modified Cholesky approximations: GMW_81, GMW_T1, GMW_T2, SE_90, SE_99, SE_T1, SE_T2.

Host-side mirror of matrix/approximation/positive_semidefinite/GMW_SE.py of the reference (same function names,
arguments and return value: the positive definite approximation B of A).  The O(n^3) part -- the modified LDL^T
without pivoting in which d_k is chosen from |c|_inf (GMW, GMW_SE.py:121-152) or |c|_1 and a 2 x 2 eigenvalue step for
the last two columns (SE, GMW_SE.py:155-194) of the current column, GMW_SE.py:6-35, :95-118 -- runs on the device
(mdb_gmw_f64 / mdb_se_f64); B = A + diag(delta) and the optional rescaling of the diagonal (GMW_SE.py:43-62) are
O(n^2) numpy operations here.

The reference's own SE_* functions raise ValueError with numpy >= 1.25 (GMW_SE.py:180 builds
np.array([[a, c], [c, A]]) from a length-1 vector c and a 1 x 1 matrix A, which older numpy versions converted to
scalars); the golden vectors these functions are tested against were generated from the reference with exactly that
conversion restored (tests/golden/generate_se_goldens.py).
"""
import math

import numpy as np

from ... import _native, errors

__all__ = ['GMW_81', 'GMW_T1', 'GMW_T2', 'SE_90', 'SE_99', 'SE_T1', 'SE_T2']


def _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases, nondecreasing_startegy, use_abs,
         revisited_version_mu, family='gmw'):
    A = np.asarray(A)                                                   # GMW_SE.py:69-72
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise errors.MatrixNotSquareError(A)
    if np.iscomplexobj(A):
        raise ValueError('complex matrices are outside the dense real (float64) path of this package.')
    if min_diag_D is None:                                              # :76-80
        min_diag_D = float(np.finfo(np.float64).eps) ** 0.5
    if not min_diag_D > 0:
        raise ValueError('min_diag_D must be positive.')                # the reference asserts it (:81)
    n = A.shape[0]
    A64 = np.ascontiguousarray(A, dtype=np.float64)
    d = np.empty(n)
    delta = np.empty(n)
    ctx = _native.context()
    mu = math.nan if revisited_version_mu is None else float(revisited_version_mu)
    fn = ctx.lib.mdb_se_f64 if family == 'se' else ctx.lib.mdb_gmw_f64
    rc = fn(ctx.handle, n, _native.ptr(A64), n, _native.HOST, int(use_two_phases), int(nondecreasing_startegy),
            int(use_abs), mu, float(min_diag_D), _native.ptr(d), _native.ptr(delta))
    ctx.check(rc, 'mdb_se_f64' if family == 'se' else 'mdb_gmw_f64')
    # _modified_A, GMW_SE.py:43-62
    B = A if overwrite_A and A.dtype == np.float64 else A64.copy()
    B[np.diag_indices(n)] += delta
    if min_diag_B is not None or max_diag_B is not None:
        lo = -np.inf if min_diag_B is None else min_diag_B
        hi = np.inf if max_diag_B is None else max_diag_B
        diag_old = B.diagonal().copy()
        diag_new = np.minimum(np.maximum(diag_old, lo), hi)
        s = (diag_new / diag_old) ** 0.5
        B *= s[:, np.newaxis]
        B *= s[np.newaxis, :]
        B[np.diag_indices(n)] = diag_new
    return B


def GMW_81(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, overwrite_A=False):
    """Positive definite approximation of `A` by the algorithm of Gill, Murray and Wright 1981 (GMW_SE.py:197-243)."""
    return _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases=False, nondecreasing_startegy=False,
                use_abs=True, revisited_version_mu=None)


def GMW_T1(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, overwrite_A=False):
    """GMW-I of Fang and O'Leary 2008 (GMW_SE.py:246-293)."""
    return _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases=True, nondecreasing_startegy=True,
                use_abs=True, revisited_version_mu=0.75)


def GMW_T2(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, overwrite_A=False):
    """GMW-II of Fang and O'Leary 2008 (GMW_SE.py:296-343)."""
    return _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases=True, nondecreasing_startegy=True,
                use_abs=False, revisited_version_mu=0.75)


def SE_90(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, overwrite_A=False):
    """Positive definite approximation of `A` by the algorithm of Schnabel and Eskow 1990 (GMW_SE.py:374-430)."""
    return _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases=True, nondecreasing_startegy=True,
                use_abs=False, revisited_version_mu=None, family='se')


def SE_99(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, overwrite_A=False):
    """The revised algorithm of Schnabel and Eskow 1999 (GMW_SE.py:433-493)."""
    return _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases=True, nondecreasing_startegy=True,
                use_abs=False, revisited_version_mu=0.1, family='se')


def SE_T1(A, min_diag_B=None, max_diag_B=None, min_diag_D=None, overwrite_A=False):
    """SE-I of Fang and O'Leary 2008 (GMW_SE.py:496-555)."""
    return _gmw(A, min_diag_B, max_diag_B, min_diag_D, overwrite_A, use_two_phases=True, nondecreasing_startegy=False,
                use_abs=True, revisited_version_mu=0.1, family='se')


SE_T2 = SE_99   # GMW_SE.py:558
