"""Approximation of symmetric matrices by positive semidefinite matrices
(reference: matrix/approximation/positive_semidefinite/__init__.py)."""
from .reimer import (decomposition, positive_semidefinite_matrix,
                     APPROXIMATION_ONLY_PERMUTATION_METHODS, MAXIMAL_STABILITY_PERMUTATION_METHOD,
                     MINIMAL_DIFFERENCE_PERMUTATION_METHOD)
