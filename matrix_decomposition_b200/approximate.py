"""Deprecated alias module kept for callers of `matrix.approximate` (reference: matrix/approximate.py:1-15): the
functions live in `approximation.positive_semidefinite`; importing this module warns once."""
import warnings as _warnings

from .approximation import positive_semidefinite as _psd

_warnings.warn('matrix.approximate is deprecated and will be removed in a future release; use '
               'matrix.approximation.positive_semidefinite instead.', DeprecationWarning, stacklevel=2)

APPROXIMATION_ONLY_PERMUTATION_METHODS = _psd.APPROXIMATION_ONLY_PERMUTATION_METHODS
decomposition = _psd.decomposition
positive_definite_matrix = _psd.positive_semidefinite_matrix     # the old name of today's positive_semidefinite_matrix


def positive_semidefinite_matrix(*args, **kwargs):
    """ The old `positive_semidefinite_matrix`: D may reach zero (approximate.py:14-15). """
    return _psd.positive_semidefinite_matrix(*args, min_diag_D=0, **kwargs)
