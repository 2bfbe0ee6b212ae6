"""
This module is deprecated and will be removed in a future release.
Please use matrix.approximation.positive_semidefinite instead.  (reference: matrix/approximate.py:1-15)
"""

import warnings

warnings.warn("This module is deprecated and will be removed in a future release. Please use "
              "matrix.approximation.positive_semidefinite instead.", DeprecationWarning, stacklevel=2)

from .approximation.positive_semidefinite import decomposition, APPROXIMATION_ONLY_PERMUTATION_METHODS  # noqa: E402,F401
from .approximation.positive_semidefinite import positive_semidefinite_matrix as positive_definite_matrix  # noqa: E402


def positive_semidefinite_matrix(*args, **kargs):
    return positive_definite_matrix(*args, min_diag_D=0, **kargs)
