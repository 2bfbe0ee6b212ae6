"""Synthetic input families for tests, golden generation and bench.py.

All float64, seeded, host-side numpy.  The device-side generator used above the oracle's range
(n = 65536 / 131072) lives in csrc (mdb_generate_f64) and follows family F2.
"""
import math

import numpy as np


def goe(n, seed):
    """F1: symmetric Gaussian matrix A = (G + G^T) / 2, G ~ N(0, 1)."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    return (G + G.T) / 2


def shifted_goe(n, seed, s=1.05):
    """F2: A = GOE + mu I with mu = s sqrt(2 n) and the mu-scaled bounds of SURVEY 8d.

    s = 1.05 is the benign (SPD, upper-bound clamps only) regime used for the headline benchmark;
    s = 0.9 is slightly indefinite (explosive regime, compare clamped set and B only).
    Returns (A, kwargs for approximation.positive_semidefinite.decomposition).
    """
    mu = s * math.sqrt(2 * n)
    A = goe(n, seed) + mu * np.eye(n)
    return A, f2_bounds(n, s)


def f2_bounds(n, s=1.05):
    mu = s * math.sqrt(2 * n)
    return dict(min_diag_D=0.015 * mu, min_abs_value_D=0.015 * mu, max_diag_D=0.9 * mu, max_diag_B=2 * mu,
                permutation='decreasing_diagonal_values')


def lowrank_plus_diag(n, seed, lo=0.5, hi=1.5, k=64):
    """F3: A = W W^T / k + diag(u), W ~ N(0,1)^{n x k}, u ~ U(lo, hi).  (0.5, 1.5) is SPD."""
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((n, k))
    u = rng.uniform(lo, hi, n)
    A = W @ W.T / k
    A[np.diag_indices(n)] += u
    return A


def reference_fixture(n):
    """F4: matrix.tests.random.hermitian_matrix(n) * 10 (tests/random.py:18-23,82-107, dense real):
    legacy np.random.seed(1234); U = rand(n, n); (U + U^T) * 10."""
    state = np.random.get_state()
    try:
        np.random.seed(1234)
        U = np.random.rand(n, n)
    finally:
        np.random.set_state(state)
    return (U + U.T) * 10


def reference_fixture_spd(n):
    """matrix.tests.random.hermitian_matrix(n, positive_semidefinite=True, invertible=True)
    (tests/random.py:90-107): L D L^T with unit-lower L from rand, d = rand + 1."""
    state = np.random.get_state()
    try:
        np.random.seed(1234)
        U = np.random.rand(n, n)
        d = np.random.rand(n) + 1
    finally:
        np.random.set_state(state)
    L = np.tril(U, -1)
    L[np.diag_indices(n)] = 1
    return L @ np.diag(d) @ L.T


def reference_permutation(n):
    """matrix.tests.random.permutation_vector (tests/random.py:134-136)."""
    state = np.random.get_state()
    try:
        np.random.seed(1234)
        return np.random.permutation(n)
    finally:
        np.random.set_state(state)
