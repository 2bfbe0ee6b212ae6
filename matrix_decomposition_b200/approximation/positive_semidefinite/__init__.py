"""Approximation of symmetric matrices by positive semidefinite matrices
(reference: matrix/approximation/positive_semidefinite/__init__.py)."""
from .reimer import (decomposition, positive_semidefinite_matrix, decomposition_batched, shard_for_rank,
                     BatchedDecompositions,
                     APPROXIMATION_ONLY_PERMUTATION_METHODS, MAXIMAL_STABILITY_PERMUTATION_METHOD,
                     MINIMAL_DIFFERENCE_PERMUTATION_METHOD)
from .gmw_se import GMW_81, GMW_T1, GMW_T2, SE_90, SE_99, SE_T1, SE_T2  # noqa: E402,F401  (GMW_SE.py; __init__.py:14)
