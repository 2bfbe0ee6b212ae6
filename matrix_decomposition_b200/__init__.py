"""matrix_decomposition_b200 -- B200-native (sm_100a) drop-in for the dense factorization hot path of
jor-/matrix-decomposition (`import matrix_decomposition_b200 as matrix`).

Facade with the reference's names (matrix/__init__.py:3-55): decompose, is_positive_definite, ...,
the constants, and the sub-modules approximation, decompositions, errors, permute.  Compute runs in
libmdb200.so (hand-written CUDA); see DESIGN.md.  There is no CPU fallback.
"""
import logging

logger = logging.getLogger('matrix')  # the reference's logger name (matrix/__init__.py:38-40)

from . import constants, errors, util, permute, decompositions  # noqa: E402
from . import approximation  # noqa: E402
from .calculate import (is_positive_semidefinite, is_positive_definite, is_invertible, decompose, solve)  # noqa: E402
from .constants import (  # noqa: E402,F401
    DECOMPOSITION_TYPES,
    LDL_DECOMPOSITION_TYPE, LDL_DECOMPOSITION_COMPRESSED_TYPE, LL_DECOMPOSITION_TYPE,
    UNIVERSAL_PERMUTATION_METHODS, SPARSE_ONLY_PERMUTATION_METHODS,
    NO_PERMUTATION_METHOD,
    DECREASING_DIAGONAL_VALUES_PERMUTATION_METHOD, INCREASING_DIAGONAL_VALUES_PERMUTATION_METHOD,
    DECREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD,
    INCREASING_ABSOLUTE_DIAGONAL_VALUES_PERMUTATION_METHOD)

__version__ = '0.1.0'


def release_device_memory():
    """Not in the reference: returns the device buffers that libmdb200 keeps between calls (the working matrix and
    the upload staging area of the last large factorization, mdb_ctx_trim in include/mdb200.h) to the CUDA driver,
    e.g. before another library in the same process needs the memory."""
    from . import _native
    for ctx in list(_native._contexts.values()):
        ctx.trim()


def __getattr__(name):
    """Deprecated aliases of the reference (matrix/__init__.py:45-55)."""
    deprecated_names = ['decomposition', 'positive_definite_matrix', 'positive_semidefinite_matrix',
                        'APPROXIMATION_ONLY_PERMUTATION_METHODS']
    if name in deprecated_names:
        import warnings
        warnings.warn(f'"matrix.{name}" is deprecated. Take a look at'
                      ' "matrix.approximation.positive_semidefinite" instead.', DeprecationWarning, stacklevel=2)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', DeprecationWarning)
            from . import approximate
        return getattr(approximate, name)
    raise AttributeError(f'Module {__name__} has no attribute {name}.')
