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


def _reject(message, cause=None):
    """Logs and raises the ValueError every argument check of R.py:291-457 ends in."""
    error = ValueError(message)
    logger.error(error)
    if cause is not None:
        raise error from cause
    raise error


def _infinity_policy(allow_minus, allow_plus):
    """(predicate on an ndarray, description) for the values a bound may take besides finite ones."""
    extras = [(-math.inf, 'minus infinity')] * allow_minus + [(math.inf, 'plus infinity')] * allow_plus

    def admissible(values):
        mask = np.isfinite(values)
        for special, _ in extras:
            mask = mask | (values == special)
        return mask
    return admissible, ''.join(' or ' + text for _, text in extras)


def _diag_bound(value, n, name, default, allow_minus=False, allow_plus=False):
    """A bound on diag(B): None -> default, else a real scalar or an n-vector (contract of R.py:291-322)."""
    if value is None:
        return np.asarray(default)
    try:
        arr = np.asarray(value)
    except TypeError as cause:
        _reject(f'{name}: expected a scalar or a vector, got {value!r}.', cause)
    if arr.shape not in ((), (n,)):
        _reject(f'{name}: expected a scalar or a vector of length {n} (the dimension of A), got shape {arr.shape}.')
    if not np.all(np.isreal(arr)):
        _reject(f'{name}: expected real values, got {arr}.')
    admissible, extra_text = _infinity_policy(allow_minus, allow_plus)
    if not np.all(admissible(arr)):
        _reject(f'{name}: every value has to be finite{extra_text}, got {arr}.')
    return arr


def _scalar_bound(value, name, default=None, at_least=None, at_most=None, allow_plus=False):
    """A scalar option: None -> default, else a float inside [at_least, at_most] (contract of R.py:324-354)."""
    if value is None:
        return default
    try:
        x = float(value)
    except TypeError as cause:
        _reject(f'{name}: expected a float, got {value!r}.', cause)
    if not (math.isfinite(x) or (allow_plus and x == math.inf)):
        _reject(f'{name}: has to be finite{" or plus infinity" if allow_plus else ""}, got {x}.')
    if at_least is not None and x < at_least:
        _reject(f'{name}: has to be >= {at_least}, got {x}.')
    if at_most is not None and x > at_most:
        _reject(f'{name}: has to be <= {at_most}, got {x}.')
    return x


_DYNAMIC_CODES = {MINIMAL_DIFFERENCE_PERMUTATION_METHOD: 1, MAXIMAL_STABILITY_PERMUTATION_METHOD: 2}


def _resolve_permutation(A, n, permutation, min_diag_D):
    """-> (p or None, dynamic code).  Strings name a static method (resolved on the host exactly like
    permute.permutation_vector) or one of the two dynamic pivoting rules (resolved on the device, starting from the
    identity, R.py:443); anything else must be a length-n vector (R.py:414-457).  None = 'maximal_stability'."""
    if permutation is None:
        permutation = MAXIMAL_STABILITY_PERMUTATION_METHOD
    if not isinstance(permutation, str):
        p = np.asanyarray(permutation)
        if p.shape != (n,):
            _reject(f'permutation: a vector must have one entry per row of A ({n}), got shape {p.shape}.')
        return p, 0
    method = permutation.lower()
    if method in _DYNAMIC_CODES:
        if method == MINIMAL_DIFFERENCE_PERMUTATION_METHOD and min_diag_D is not None and min_diag_D <= 0:
            _reject(f"permutation '{method}' needs min_diag_D > 0 (R.py:437-441).")
        return None, _DYNAMIC_CODES[method]
    if method not in constants.UNIVERSAL_PERMUTATION_METHODS:
        known = constants.UNIVERSAL_PERMUTATION_METHODS + APPROXIMATION_ONLY_PERMUTATION_METHODS
        _reject(f"permutation '{method}' is not one of {known}.")
    return permute.permutation_vector(A, permutation_method=method), 0


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

    min_diag_B = _diag_bound(min_diag_B, n, 'min_diag_B', -math.inf, allow_minus=True)
    max_diag_B = _diag_bound(max_diag_B, n, 'max_diag_B', math.inf, allow_plus=True)
    min_diag_D = _scalar_bound(min_diag_D, 'min_diag_D', at_least=0)
    max_diag_D = _scalar_bound(max_diag_D, 'max_diag_D', default=math.inf, allow_plus=True)
    lowest_allowed = max(np.max(min_diag_B), 0 if min_diag_D is None else min_diag_D)
    highest_allowed = min(np.min(max_diag_B), max_diag_D)
    if not lowest_allowed <= highest_allowed:  # R.py:368-373
        _reject(f'the lower bounds (min_diag_B, min_diag_D: {lowest_allowed}) exceed the upper bounds '
                f'(max_diag_B, max_diag_D: {highest_allowed}).')

    # floors of R.py:396-408: |d| >= eps(float128), thresholding of L at >= eps(float64)
    eps_D = np.finfo(np.longdouble).eps   # stays a longdouble so that its square root rounds like the reference's
    eps_L = float(np.finfo(np.float64).eps)
    min_abs_value_D = max(_scalar_bound(min_abs_value_D, 'min_abs_value_D', default=eps_D ** 0.5, at_least=0), eps_D)
    min_abs_value_L = max(_scalar_bound(min_abs_value_L, 'min_abs_value_L', default=eps_L, at_least=0, at_most=1), eps_L)

    p, dynamic = _resolve_permutation(A, n, permutation, min_diag_D)
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
    R.py:68.  `A` is never modified (`overwrite_A` only documents permission).  `A` is symmetric by contract (the
    reference's loop reads A[j, i] and A[i, j] interchangeably): the static-permutation path reads it through its upper
    triangle only, at every size, so a matrix that is not exactly symmetric gives size-independent results.

    The dynamic pivoting methods 'maximal_stability' (the default for permutation=None, R.py:415-416) and
    'minimal_difference' run as an unblocked right-looking factorization with a device-wide arg-min per
    column (mdb_approximate_dynamic_f64); they need n^2/2 rule evaluations and are meant for n up to ~10^4.
    """
    if return_type is not None and return_type not in constants.DECOMPOSITION_TYPES:
        _reject(f'return_type {return_type!r} is not one of {constants.DECOMPOSITION_TYPES}.')
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


def shard_for_rank(batch, rank=None, world=None):
    """Slice of a batch of independent matrices that rank `rank` of `world` processes factors (contiguous shares,
    sizes differ by at most one).  Defaults: RANK / WORLD_SIZE of the launcher (torchrun), else the whole batch.
    Independent units: no data-path collective (BASELINE config 5)."""
    import os
    rank = int(os.environ.get('RANK', '0')) if rank is None else int(rank)
    world = int(os.environ.get('WORLD_SIZE', '1')) if world is None else int(world)
    if not 0 <= rank < world:
        raise ValueError(f'rank {rank} is not in [0, {world}).')
    base, extra = divmod(int(batch), world)
    begin = rank * base + min(rank, extra)
    return slice(begin, begin + base + (1 if rank < extra else 0))


class BatchedDecompositions:
    """Result of `decomposition_batched`: a read-only sequence of `batch` LDL decompositions that share their storage
    (L: batch x n x n, d / omega / delta: batch x n in float64, p: batch x n, clamped: batch x n bool).
    `result[b]` builds the reference's `LDL_Decomposition` for matrix b with `.omega`, `.delta`, `.clamped` attached
    exactly as `decomposition` does (R.py:810-812); the arrays themselves are available for bulk use."""

    def __init__(self, L, d, p, omega, delta, clamped):
        self.L, self.d, self.p, self.omega, self.delta, self.clamped = L, d, p, omega, delta, clamped

    def __len__(self):
        return self.d.shape[0]

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        b = range(len(self))[b]
        dec = decompositions.LDL_Decomposition(self.L[b], self.d[b].astype(RESULT_DTYPE), p=self.p[b])
        dec.omega = self.omega[b].astype(RESULT_DTYPE)
        dec.delta = self.delta[b].astype(RESULT_DTYPE)
        dec.clamped = self.clamped[b]
        return dec

    def __iter__(self):
        return (self[b] for b in range(len(self)))


def decomposition_batched(As, min_diag_B=None, max_diag_B=None, min_diag_D=None, max_diag_D=None,
                          min_abs_value_D=None, min_abs_value_L=None, permutation=None):
    """
    `decomposition(A, ...)` for a stack of independent matrices of one size in ONE device pass: `As` has shape
    (batch, n, n) (or is a sequence of n x n matrices); the options mean what they mean for `decomposition`
    (R.py:712-816) and apply to every matrix (a vector `min_diag_B` / `max_diag_B` is indexed by the original row
    index of each matrix).  Returns a `BatchedDecompositions`; element b equals `decomposition(As[b], ...,
    return_type='LDL')`.  This is the many-small-matrix regime (BASELINE config 5: 4096 matrices of 512 x 512):
    the batch rides on the grid's z dimension of every kernel (mdb_approximate_batched_f64) instead of 12 launches
    per matrix.  A process factors what it is given on its own GPU (LOCAL_RANK); shard a larger job over the ranks
    of a torchrun launch with `As[shard_for_rank(len(As))]` -- independent units, no collective.

    Static permutation methods and explicit vectors (one length-n vector for all matrices, or batch x n) run
    batched; the two dynamic pivoting rules need a device-wide arg-min per column and fall back to one
    `decomposition` call per matrix.
    """
    As = np.asarray(As)
    if As.ndim != 3 or As.shape[1] != As.shape[2]:
        raise errors.MatrixNotSquareError(As)
    batch, n = As.shape[0], As.shape[1]
    explicit = permutation is not None and not isinstance(permutation, str)
    if explicit:
        P_all = np.asanyarray(permutation)
        if P_all.shape not in ((n,), (batch, n)):
            _reject(f'permutation: expected one vector of length {n} or an array of shape {(batch, n)}, '
                    f'got shape {P_all.shape}.')
        P_all = np.broadcast_to(P_all, (batch, n))
    common = dict(min_diag_B=min_diag_B, max_diag_B=max_diag_B, min_diag_D=min_diag_D, max_diag_D=max_diag_D,
                  min_abs_value_D=min_abs_value_D, min_abs_value_L=min_abs_value_L)
    if batch == 0:
        z = np.zeros((0, n))
        return BatchedDecompositions(np.zeros((0, n, n)), z, np.zeros((0, n), dtype=np.int64), z, z, np.zeros((0, n), dtype=bool))
    # the shared options are validated once, against the first matrix (R.py:273-457)
    first = _prepare(As[0], permutation=P_all[0] if explicit else permutation, **common)
    _, _, _, dynamic, mB, MB, mD, MD, mavD, mavL = first
    if dynamic:
        decs = [decomposition(As[b], permutation=permutation, **common) for b in range(batch)]
        return BatchedDecompositions(np.stack([x.L for x in decs]), np.stack([np.asarray(x.d, dtype=np.float64) for x in decs]),
                                     np.stack([np.asarray(x.p, dtype=np.int64) for x in decs]),
                                     np.stack([np.asarray(x.omega, dtype=np.float64) for x in decs]),
                                     np.stack([np.asarray(x.delta, dtype=np.float64) for x in decs]),
                                     np.stack([x.clamped for x in decs]))
    if np.issubdtype(As.dtype, np.complexfloating):
        raise ValueError('complex matrices are outside the dense real (float64) path of this package.')
    A64 = np.ascontiguousarray(As, dtype=np.float64)
    if explicit:
        ps = np.ascontiguousarray(P_all, dtype=np.int64)
    else:
        method = (permutation or MAXIMAL_STABILITY_PERMUTATION_METHOD).lower()
        ps = np.stack([np.asarray(permute.permutation_vector(A64[b], permutation_method=method), dtype=np.int64)
                       for b in range(batch)])
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, mD, MD, mavD, mavL, mB, MB)
    L = np.empty((batch, n, n), dtype=np.float64)
    d, omega, delta = (np.empty((batch, n), dtype=np.float64) for _ in range(3))
    clamped = np.zeros((batch, n), dtype=np.uint8)
    ctx = _native.context()
    P = _native.ptr
    rc = ctx.lib.mdb_approximate_batched_f64(ctx.handle, batch, n, P(A64), n, P(ps), ctypes.byref(bounds), P(L), n, P(d),
                                             P(omega), P(delta), P(clamped), _native.HOST)
    ctx.check(rc, 'mdb_approximate_batched_f64')
    del keep
    return BatchedDecompositions(L, d, ps, omega, delta, clamped.astype(bool))


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
