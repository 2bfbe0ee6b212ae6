"""GPU parity tests (pytest -m gpu, on the B200 box).  Everything goes through the public Python API,
i.e. through the C ABI of libmdb200.so, and is compared with the CPU oracle on the same seeded inputs,
with the golden vectors produced by the unmodified reference, and -- at BASELINE's full size -- with the
size-independent identity of Reimer.py:947-957.  /root/reference is never read here.

Tolerances (BASELINE.json north_star): permutation vector and clamped-index set bit-exact; L, d, omega,
solve results within 1e-10 (|dL| <= 1e-10 max(1, |L|), relative for d / x) in the benign regime;
||B - B_ref||_F / ||B_ref||_F <= 1e-12 for the composed matrix B = L D L^T in every regime (in the
explosive regime d / L individually are decided by rounding noise in the reference itself).
"""
import ctypes
import json
import math
import os
import sys

import numpy as np
import pytest

import workloads
from conftest import load_golden, ld_join

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL = 1e-10
TOL_B = 1e-12


@pytest.fixture(scope='module')
def matrix():
    import matrix_decomposition_b200 as m
    from matrix_decomposition_b200 import _native
    _native.context()          # fails loudly without a GPU / libmdb200.so
    return m


@pytest.fixture(scope='module')
def orc():
    from oracle import cpu_oracle
    return cpu_oracle


def _kw_from_json(kw):
    out = {}
    for k, v in kw.items():
        out[k] = np.asarray(v, dtype=float) if isinstance(v, list) and k != 'permutation' else v
        if k == 'permutation' and isinstance(v, list):
            out[k] = np.asarray(v, dtype=np.int64)
    return out


def _composed(L, d):
    d = np.asarray(d, dtype=np.float64)
    return L @ (d[:, None] * L.T)


def _compare(dec, ref_L, ref_d, ref_p, ref_clamped, ref_omega=None, ref_delta=None, benign=True, scale=1.0):
    np.testing.assert_array_equal(np.asarray(dec.p), ref_p)
    np.testing.assert_array_equal(dec.clamped, ref_clamped)
    L, d = dec.L, np.asarray(dec.d, dtype=np.float64)
    ref_d = np.asarray(ref_d, dtype=np.float64)
    B, B_ref = _composed(L, d), _composed(ref_L, ref_d)
    assert np.linalg.norm(B - B_ref) <= TOL_B * np.linalg.norm(B_ref)
    if benign:
        assert np.all(np.abs(L - ref_L) <= TOL * np.maximum(1.0, np.abs(ref_L)))
        np.testing.assert_allclose(d, ref_d, rtol=TOL)
        if ref_omega is not None:
            np.testing.assert_allclose(np.asarray(dec.omega, dtype=float), np.asarray(ref_omega, dtype=float), rtol=TOL,
                                       atol=1e-14)
        if ref_delta is not None:
            np.testing.assert_allclose(np.asarray(dec.delta, dtype=float), np.asarray(ref_delta, dtype=float),
                                       rtol=1e-9, atol=1e-11 * scale)


# ---------------------------------------------------------------------------------------------
# approximate: golden vectors of the reference
# ---------------------------------------------------------------------------------------------
def test_known_answer_2x2(matrix):
    dec = matrix.approximation.positive_semidefinite.decomposition(
        np.array([[2.0, 1.0], [1.0, -3.0]]), min_diag_D=1e-6, permutation='decreasing_diagonal_values')
    np.testing.assert_array_equal(dec.p, [0, 1])
    np.testing.assert_allclose(np.asarray(dec.d, dtype=float), [2, 1e-6], rtol=1e-15)
    np.testing.assert_allclose(np.asarray(dec.omega, dtype=float), [1, 0.3938888], rtol=1e-6)
    np.testing.assert_allclose(np.asarray(dec.delta, dtype=float), [0, 3.07757519], rtol=1e-7, atol=1e-15)
    np.testing.assert_allclose(dec.L[1, 0], 0.1969444, rtol=1e-6)
    assert list(dec.clamped) == [False, True]


def test_reference_test_grid_n10_static_permutations(matrix):
    """Every static-permutation parametrisation of matrix/tests/test_approximate.py (dense, real) that the
    reference completes on its own fixture: 500+ cases against the reference's own outputs."""
    z = load_golden('approx_fixture10.npz')
    A = z['A']
    manifest = json.loads(str(z['manifest']))
    n_run = n_ties = 0
    for m in manifest:
        kw = _kw_from_json(m['kwargs'])
        perm = kw['permutation']
        if not m['ok'] or (isinstance(perm, str) and perm in ('minimal_difference', 'maximal_stability')) or perm is None:
            continue        # dynamic pivoting: test_dynamic_pivoting_reference_test_grid_n10
        key = m['key']
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        d_ref = ld_join(z, key + '/d').astype(np.float64)
        np.testing.assert_array_equal(np.asarray(dec.p), z[key + '/p'])
        np.testing.assert_array_equal(dec.clamped, z[key + '/clamped'])
        B, B_ref = _composed(dec.L, dec.d), _composed(z[key + '/L'], d_ref)
        assert np.linalg.norm(B - B_ref) <= TOL_B * np.linalg.norm(B_ref), key
        d = np.asarray(dec.d, dtype=np.float64)
        if np.all(np.abs(d - d_ref) <= 1e-9 * np.abs(d_ref)):
            assert np.all(np.abs(dec.L - z[key + '/L']) <= 1e-9 * np.maximum(1.0, np.abs(z[key + '/L']))), key
        else:
            n_ties += 1          # degenerate (d, omega) tie of SURVEY fact 7: B and the clamped set still agree
        n_run += 1
    assert n_run >= 500
    assert n_ties <= 0.10 * n_run   # FP64 cannot resolve f-ties that float128 resolves by rounding noise (class iii)


@pytest.mark.parametrize('key,benign', [('goe64', False), ('f2_n200_s1.05', True), ('f2_n200_s0.9', False),
                                        ('f3_n128_nearpsd', True), ('f3_n128_nearpsd_default', True)])
def test_families_against_reference_goldens(matrix, key, benign):
    z = load_golden('approx_families.npz')
    m = {e['key']: e for e in json.loads(str(z['manifest']))}[key]
    fam = m['family']
    if fam['kind'] == 'goe':
        A = workloads.goe(fam['n'], fam['seed'])
    elif fam['kind'] == 'shifted_goe':
        A = workloads.shifted_goe(fam['n'], fam['seed'], fam['s'])[0]
    else:
        A = workloads.lowrank_plus_diag(fam['n'], fam['seed'], fam['lo'], fam['hi'])
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **_kw_from_json(m['kwargs']))
    _compare(dec, z[key + '/L'], ld_join(z, key + '/d'), z[key + '/p'], z[key + '/clamped'],
             ld_join(z, key + '/omega'), ld_join(z, key + '/delta'), benign=benign, scale=np.abs(A).max())


@pytest.mark.parametrize('key,benign', [('f2_n1000_s1.05', True), ('f2_n600_s0.9', False), ('c1_goe2000', False)])
def test_large_goldens_decisions_and_probes(matrix, key, benign):
    """BASELINE config 1 (GOE n=2000, min_diag_D=1e-6) and two F2 cases: decisions and B v probes of the reference."""
    z = load_golden('approx_families.npz')
    m = {e['key']: e for e in json.loads(str(z['manifest']))}[key]
    fam = m['family']
    A = workloads.goe(fam['n'], fam['seed']) if fam['kind'] == 'goe' else workloads.shifted_goe(fam['n'], fam['seed'], fam['s'])[0]
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **_kw_from_json(m['kwargs']))
    np.testing.assert_array_equal(np.asarray(dec.p), z[key + '/p'])
    np.testing.assert_array_equal(dec.clamped, z[key + '/clamped'])
    d = np.asarray(dec.d, dtype=np.float64)
    V = z[key + '/probes']
    Bv = dec.L @ (d[:, None] * (dec.L.T @ V))
    assert np.max(np.abs(Bv - z[key + '/Bv'])) <= 1e-11 * np.max(np.abs(z[key + '/Bv']))
    if benign:
        np.testing.assert_allclose(d, ld_join(z, key + '/d').astype(float), rtol=TOL)
        rows = z[key + '/L_rows_idx']
        assert np.all(np.abs(dec.L[rows] - z[key + '/L_rows']) <= TOL * np.maximum(1.0, np.abs(z[key + '/L_rows'])))
    # solve with the resident factor reproduces A_approx x = b
    b = np.random.default_rng(1).standard_normal(A.shape[0])
    x = dec.solve(b)
    r = dec.matrix_right_side_multiplication(x) - b
    if benign:
        assert np.max(np.abs(r)) <= 1e-8 * max(1.0, np.max(np.abs(b)))
    else:
        # explosive regime (|L| up to 1e60): the honest bound is the componentwise backward error of a triangular
        # solve pair, eps-level relative to |L| |D| |L^T| |x| (+ |b|); it also pins the pre-inverted 128 x 128
        # diagonal blocks of the solve where they are least stable
        Labs = np.abs(dec.L)
        scale = Labs @ (np.abs(d) * (Labs.T @ np.abs(x)))
        assert np.all(np.isfinite(x))
        assert np.max(np.abs(r) / (scale + np.abs(b))) <= 1e-7     # measured 5.5e-9 (c1_goe2000), 1e-16-level when benign


# ---------------------------------------------------------------------------------------------
# approximate: oracle on seeded inputs, edge shapes, options
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 3, 127, 128, 129, 255, 257, 384, 700])
def test_sizes_around_block_boundaries(matrix, orc, n):
    A, kw = workloads.shifted_goe(n, 100 + n, 1.05)
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    ref = orc.approximate(A, **kw)
    _compare(dec, ref['L'], ref['d'], ref['p'], ref['clamped'], ref['omega'], ref['delta'], scale=np.abs(A).max())
    assert np.all(np.triu(dec.L, 1) == 0) and np.all(dec.L.diagonal() == 1)


@pytest.mark.parametrize('blocking', [dict(fixed=2), dict(fixed=4), dict(fixed=8), dict(fixed=0, t2=300, t4=600, t8=1100)])
@pytest.mark.parametrize('n', [640, 1500])
def test_outer_block_aggregation_matches_oracle(matrix, orc, n, blocking):
    """Two-level blocking (bulk trailing update per outer block of 2 / 4 / 8 panels, shrinking with the remaining
    size) is what large n uses by default; forced here at sizes the oracle finishes in seconds."""
    from matrix_decomposition_b200 import _native
    ctx = _native.context()
    A, kw = workloads.shifted_goe(n, 900 + n, 1.05)
    ref = orc.approximate(A, **kw)
    try:
        ctx.set_blocking(**blocking)
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        S = workloads.lowrank_plus_diag(n, 5)
        dd = matrix.decompose(S, permutation='decreasing_diagonal_values', return_type='LL')
    finally:
        ctx.set_blocking()
    _compare(dec, ref['L'], ref['d'], ref['p'], ref['clamped'], ref['omega'], ref['delta'], scale=np.abs(A).max())
    rr = orc.decompose(S, 'decreasing_diagonal_values', 'LL')
    np.testing.assert_allclose(dd.L, rr['L'], rtol=1e-10, atol=1e-12)


def test_randomized_configurations_against_oracle(matrix, orc):
    """Seeded sweep over bound combinations the fixed cases do not hit: scalar / vector / absent B bounds, min_diag_D
    None / 0 / positive, finite and infinite max_diag_D, all static permutations, sizes across two to four panels.
    Decisions (permutation, clamped set) must equal the oracle's, the composed matrix must agree to 1e-11."""
    rng = np.random.default_rng(20261020)
    perms = ['none', 'decreasing_diagonal_values', 'increasing_diagonal_values', 'decreasing_absolute_diagonal_values',
             'increasing_absolute_diagonal_values']
    n_checked = 0
    for trial in range(24):
        n = int(rng.choice([130, 200, 260, 390, 515]))
        shift = float(rng.choice([0.6, 1.0, 1.3]))
        A = workloads.goe(n, 500 + trial) + shift * np.sqrt(2 * n) * np.eye(n)
        mu = np.sqrt(2 * n)
        kw = dict(permutation=perms[trial % len(perms)])
        mode = trial % 4
        if mode == 0:      # upper clamps only (benign)
            kw.update(min_diag_D=0.02 * mu, max_diag_D=0.8 * mu * shift, max_diag_B=2.5 * mu * shift)
        elif mode == 1:    # vector B bounds, default min_diag_D
            kw.update(min_diag_B=np.full(n, 0.05 * mu) + rng.random(n), max_diag_B=np.full(n, 3.0 * mu * shift) + rng.random(n))
        elif mode == 2:    # lower clamp with min_abs_value_D, no upper bounds
            kw.update(min_diag_D=0.3 * mu, min_abs_value_D=0.3 * mu)
        else:              # min_diag_D = 0: d may become exactly 0
            kw.update(min_diag_D=0.0, max_diag_D=1.2 * mu * shift, min_abs_value_D=1e-3)
        ref = orc.approximate(A, **kw)
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        np.testing.assert_array_equal(np.asarray(dec.p), ref['p'], err_msg=str(trial))
        B, B_ref = _composed(dec.L, dec.d), _composed(ref['L'], ref['d'])
        assert np.linalg.norm(B - B_ref) <= 1e-11 * np.linalg.norm(B_ref), (trial, kw)
        np.testing.assert_array_equal(dec.clamped, ref['clamped'], err_msg=str(trial))
        n_checked += 1
    assert n_checked == 24


def test_host_matrix_is_read_through_its_upper_triangle(matrix):
    """symmetry is a contract, not something that is checked -- so that a matrix that is not exactly symmetric
    gives the same result at every size, a host matrix is read through its upper triangle only (above 4096 rows nothing
    else is uploaded).  Garbage below the diagonal must not change anything, for one matrix and for a batch."""
    n = 300
    A, kw = workloads.shifted_goe(n, 11, 1.05)
    G = A.copy()
    G[np.tril_indices(n, -1)] = np.random.default_rng(0).standard_normal(n * (n - 1) // 2)
    psd = matrix.approximation.positive_semidefinite
    a, g = psd.decomposition(A, **kw), psd.decomposition(G, **kw)
    assert np.array_equal(a.p, g.p) and np.array_equal(a.clamped, g.clamped)
    assert np.array_equal(np.asarray(a.d, dtype=float), np.asarray(g.d, dtype=float)) and np.array_equal(a.L, g.L)
    b1 = psd.decomposition_batched(np.stack([A, A]), **kw)
    b2 = psd.decomposition_batched(np.stack([A, G]), **kw)
    assert np.array_equal(b1[1].L, b2[1].L) and np.array_equal(b1[0].L, a.L)


def test_empty_matrix(matrix):
    dec = matrix.approximation.positive_semidefinite.decomposition(np.zeros((0, 0)), permutation='none')
    assert dec.n == 0 and dec.L.shape == (0, 0)


@pytest.mark.parametrize('perm', ['none', 'decreasing_diagonal_values', 'increasing_diagonal_values',
                                  'decreasing_absolute_diagonal_values', 'increasing_absolute_diagonal_values', 'vector'])
@pytest.mark.parametrize('rt', [None, 'LDL', 'LDL_compressed', 'LL'])
def test_permutations_and_return_types(matrix, orc, perm, rt):
    n = 150
    A = workloads.goe(n, 77) + 3.0 * np.eye(n)
    p_arg = np.random.default_rng(5).permutation(n) if perm == 'vector' else perm
    kw = dict(min_diag_D=0.5, max_diag_D=40.0, max_diag_B=60.0, permutation=p_arg)
    A_copy = A.copy()
    dec = matrix.approximation.positive_semidefinite.decomposition(A, return_type=rt, **kw)
    assert np.array_equal(A, A_copy)                                  # A untouched (test_approximate.py:68)
    assert dec.type_str == (rt or 'LDL')
    ref = orc.approximate(A, **kw)
    ldl = dec.as_type('LDL')
    np.testing.assert_array_equal(np.asarray(dec.p), ref['p'])
    np.testing.assert_array_equal(dec.clamped, ref['clamped'])
    B, B_ref = _composed(ldl.L, ldl.d), _composed(ref['L'], ref['d'])
    assert np.linalg.norm(B - B_ref) <= TOL_B * np.linalg.norm(B_ref)
    assert np.all(np.abs(ldl.L - ref['L']) <= 1e-9 * np.maximum(1.0, np.abs(ref['L'])))
    # value checks of test_approximate.py:187-250
    d = np.asarray(ldl.d, dtype=float)
    assert np.all(np.isfinite(d)) and np.all(d >= 0.5 * (1 - 1e-12)) and np.all(d <= 40.0 * (1 + 1e-12))
    assert np.all(B.diagonal() <= 60.0 * (1 + 1e-10))
    # omega / delta identity of test_approximate.py:476-506 on the composed matrix
    Bfull = ldl.composed_matrix
    np.testing.assert_allclose(Bfull.diagonal(), A.diagonal() + np.asarray(dec.delta, dtype=float), rtol=1e-9, atol=1e-9)
    q = np.asarray(dec.p)
    om = np.asarray(dec.omega, dtype=float)
    i, j = 7, 3
    np.testing.assert_allclose(Bfull[q[i], q[j]], A[q[i], q[j]] * om[q[i]], rtol=1e-8, atol=1e-10)


def test_vector_bounds_and_default_min_diag_D(matrix, orc):
    n = 96
    A = workloads.reference_fixture(n)
    kw = dict(min_diag_D=None, min_diag_B=np.arange(n) / n, max_diag_B=np.arange(n) + 30.0, permutation='none')
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    ref = orc.approximate(A, **kw)
    _compare(dec, ref['L'], ref['d'], ref['p'], ref['clamped'], benign=False)
    B = _composed(dec.L, dec.d)
    assert np.all(B.diagonal() >= kw['min_diag_B'] - 1e-9) and np.all(B.diagonal() <= kw['max_diag_B'] + 1e-9)


def test_explosive_regime_clamped_set_and_B(matrix, orc):
    """min_diag_D = 1e-6 on a strongly indefinite matrix: Lu grows to 1e60; clamped set and B must still match."""
    A = workloads.goe(500, 9)
    kw = dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values')
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    ref = orc.approximate(A, **kw)
    _compare(dec, ref['L'], ref['d'], ref['p'], ref['clamped'], benign=False)
    assert np.all(np.isfinite(dec.L))


def test_positive_semidefinite_matrix(matrix, orc):
    z = load_golden('decompose_solve_permute.npz')
    manifest = json.loads(str(z['manifest']))
    A = workloads.reference_fixture(10)
    n_run = 0
    for m in manifest:
        if not m['key'].startswith('psd10/'):
            continue
        kw = _kw_from_json(m['kwargs'])
        B = matrix.approximation.positive_semidefinite.positive_semidefinite_matrix(A, **kw)
        np.testing.assert_allclose(B, z[m['key'] + '/B'], rtol=1e-10, atol=1e-10)
        n_run += 1
    assert n_run >= 3
    A, kw = workloads.shifted_goe(200, 8, 1.05)
    B = matrix.approximation.positive_semidefinite.positive_semidefinite_matrix(A, **kw)
    ref = orc.approximate(A, **kw)
    np.testing.assert_allclose(B, orc.psd_matrix(A, ref, None, kw['max_diag_B']), rtol=1e-10, atol=1e-10)


def test_batched_many_small_matrices(matrix, orc):
    """BASELINE config 5 shape (independent 512 x 512 problems), a small batch through the batched C-ABI call."""
    from matrix_decomposition_b200 import _native
    ctx = _native.context()
    batch, n = 6, 512
    As = np.stack([workloads.shifted_goe(n, 1000 + b, 1.05)[0] for b in range(batch)])
    kw = workloads.f2_bounds(n)
    ps = np.stack([np.argsort(-As[b].diagonal()) for b in range(batch)]).astype(np.int64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    L = np.empty((batch, n, n))
    d, om, de = np.empty((batch, n)), np.empty((batch, n)), np.empty((batch, n))
    cl = np.zeros((batch, n), dtype=np.uint8)
    P = _native.ptr
    rc = ctx.lib.mdb_approximate_batched_f64(ctx.handle, batch, n, P(As), n, P(ps), ctypes.byref(bounds), P(L), n, P(d),
                                             P(om), P(de), P(cl), _native.HOST)
    ctx.check(rc, 'mdb_approximate_batched_f64')
    for b in range(batch):
        ref = orc.approximate(As[b], **kw)
        np.testing.assert_array_equal(ps[b], ref['p'])
        np.testing.assert_array_equal(cl[b].astype(bool), ref['clamped'])
        assert np.all(np.abs(L[b] - ref['L']) <= TOL * np.maximum(1.0, np.abs(ref['L'])))
        np.testing.assert_allclose(d[b], ref['d'].astype(float), rtol=TOL)
        np.testing.assert_allclose(om[b], ref['omega'].astype(float), rtol=TOL, atol=1e-14)


# ---------------------------------------------------------------------------------------------
# dynamic pivoting (Reimer.py:497-523): 'maximal_stability' (the default) and 'minimal_difference'
# ---------------------------------------------------------------------------------------------
def _check_identity_and_bounds(A, dec, kw):
    """Oracle-free properties of test_approximate.py:187-250 and :476-506."""
    ldl = dec.as_type('LDL')
    d = np.asarray(ldl.d, dtype=float)
    p = np.asarray(dec.p)
    n = A.shape[0]
    assert sorted(p.tolist()) == list(range(n))
    B = ldl.L @ (d[:, None] * ldl.L.T)                     # permuted order
    Bo = np.empty_like(B)
    Bo[np.ix_(p, p)] = B
    om, de = np.asarray(dec.omega, dtype=float), np.asarray(dec.delta, dtype=float)
    scale = max(1.0, np.abs(A).max())
    if np.all(d != 0):
        np.testing.assert_allclose(Bo.diagonal(), A.diagonal() + de, rtol=1e-9, atol=1e-9 * scale)
        q = np.empty(n, dtype=int)
        q[p] = np.arange(n)
        later = np.where(q[:, None] > q[None, :], np.arange(n)[:, None], np.arange(n)[None, :])
        want = A * om[later]
        off = ~np.eye(n, dtype=bool)
        np.testing.assert_allclose(Bo[off], want[off], rtol=1e-8, atol=1e-9 * scale)
    if kw.get('min_diag_D') is not None:
        assert np.all((d >= kw['min_diag_D'] * (1 - 1e-12)) | (d == 0))
    if kw.get('max_diag_D') is not None:
        assert np.all(d <= kw['max_diag_D'] * (1 + 1e-12))
    if kw.get('max_diag_B') is not None:
        assert np.all(Bo.diagonal() <= np.asarray(kw['max_diag_B']) * (1 + 1e-9) + 1e-9)
    if kw.get('min_diag_B') is not None:
        assert np.all(Bo.diagonal() >= np.asarray(kw['min_diag_B']) * (1 - 1e-9) - 1e-9)


@pytest.mark.parametrize('key', ['fix96_maxstab', 'fix96_mindiff'])
def test_dynamic_pivoting_against_reference_goldens(matrix, key):
    z = load_golden('approx_families.npz')
    m = {e['key']: e for e in json.loads(str(z['manifest']))}[key]
    A = workloads.reference_fixture(m['family']['n'])
    kw = _kw_from_json(m['kwargs'])
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    _check_identity_and_bounds(A, dec, kw)
    np.testing.assert_array_equal(np.asarray(dec.p), z[key + '/p'])
    np.testing.assert_array_equal(dec.clamped, z[key + '/clamped'])
    d_ref = ld_join(z, key + '/d').astype(float)
    B, B_ref = _composed(dec.L, dec.d), _composed(z[key + '/L'], d_ref)
    assert np.linalg.norm(B - B_ref) <= 1e-11 * np.linalg.norm(B_ref)


def test_dynamic_pivoting_reference_test_grid_n10(matrix):
    """The dynamic-permutation half of the reference's own n = 10 grid (incl. permutation=None -> default)."""
    z = load_golden('approx_fixture10.npz')
    A = z['A']
    manifest = json.loads(str(z['manifest']))
    n_run = n_same_p = n_ties = 0
    for m in manifest:
        kw = _kw_from_json(m['kwargs'])
        perm = kw['permutation']
        if not m['ok'] or not (isinstance(perm, str) and perm in ('minimal_difference', 'maximal_stability')):
            continue
        key = m['key']
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        _check_identity_and_bounds(A, dec, kw)
        n_run += 1
        d_ref = ld_join(z, key + '/d').astype(float)
        if np.array_equal(np.asarray(dec.p), z[key + '/p']):
            n_same_p += 1
            np.testing.assert_array_equal(dec.clamped, z[key + '/clamped'])
            B, B_ref = _composed(dec.L, dec.d), _composed(z[key + '/L'], d_ref)
            assert np.linalg.norm(B - B_ref) <= 1e-11 * np.linalg.norm(B_ref), key
        else:
            # a different pivot is only acceptable as a tie of the selection key (class iii): at the first position
            # where the permutations differ, the two chosen candidates must carry the same (d, omega, delta) up to
            # rounding -- both keys, (-d, f, omega, j) and (f, -d, omega, j), then differ in the index alone
            i = int(np.argmax(np.asarray(dec.p) != z[key + '/p']))
            q_gpu, q_ref = int(np.asarray(dec.p)[i]), int(z[key + '/p'][i])
            d_gpu = float(np.asarray(dec.d, dtype=float)[i])
            assert abs(d_gpu - d_ref[i]) <= 1e-8 * abs(d_ref[i]), (key, i, d_gpu, d_ref[i])
            w_gpu, w_ref = float(dec.omega[q_gpu]), float(ld_join(z, key + '/omega')[q_ref])
            assert abs(w_gpu - w_ref) <= 1e-8 * max(abs(w_ref), 1e-300) + 1e-12, (key, i, w_gpu, w_ref)
            e_gpu, e_ref = float(dec.delta[q_gpu]), float(ld_join(z, key + '/delta')[q_ref])
            assert abs(e_gpu - e_ref) <= 1e-8 * abs(e_ref) + 1e-8 * np.abs(A).max(), (key, i, e_gpu, e_ref)
            n_ties += 1
    assert n_run >= 150
    # pivot choices are arg-mins over keys that tie to ~1e-16 in the explosive regime (class iii); they must
    # agree with the float128 reference in the large majority of cases
    assert n_same_p >= 0.8 * n_run, (n_same_p, n_run)
    # default permutation (None) is maximal_stability, Reimer.py:415-416
    d1 = matrix.approximation.positive_semidefinite.decomposition(A, min_diag_D=1.0)
    d2 = matrix.approximation.positive_semidefinite.decomposition(A, min_diag_D=1.0, permutation='maximal_stability')
    assert np.array_equal(d1.p, d2.p) and d1.is_equal(d2)
    with pytest.raises(ValueError):
        matrix.approximation.positive_semidefinite.decomposition(A, min_diag_D=0, permutation='minimal_difference')


@pytest.mark.parametrize('n', [400, 450])      # 450: n - 2 is a multiple of the 64 rows a block of k_dyn_column owns
def test_dynamic_pivoting_against_oracle_moderate_n(matrix, orc, n):
    A = workloads.goe(n, 21) + 2.0 * np.eye(n)
    for method in ('maximal_stability', 'minimal_difference'):
        kw = dict(min_diag_D=0.3, max_diag_D=30.0, permutation=method)
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        _check_identity_and_bounds(A, dec, kw)
        ref = orc.approximate(A, **kw)
        if np.array_equal(np.asarray(dec.p), ref['p']):
            np.testing.assert_array_equal(dec.clamped, ref['clamped'])
            B, B_ref = _composed(dec.L, dec.d), _composed(ref['L'], ref['d'])
            # growth regime (Lu ~ 1e56): FP64 vs the reference's 80-bit accumulators; the left-looking FP64 twin
            # of the oracle deviates by the same 2e-10 here, so this is precision, not the restatement
            assert np.linalg.norm(B - B_ref) <= 2e-9 * np.linalg.norm(B_ref)
        else:   # first divergence must be a key tie
            i = int(np.argmax(np.asarray(dec.p) != ref['p']))
            assert abs(float(np.asarray(dec.d, dtype=float)[i]) - float(ref['d'][i])) <= 1e-8 * abs(float(ref['d'][i]))


def test_dynamic_pivoting_identity_many_blocks(matrix):
    """n = 3000: 47 row blocks plus the blocks for the positions left of the pivot in every k_dyn_column launch, 23
    replayed panel graphs, pivot swaps across block boundaries -- checked through the oracle-free identity
    P^T L D L^T P = A o Omega + diag(delta) and the bounds (test_approximate.py:187-250, :476-506), which only hold if
    p, the rows of L, omega and delta went through the same swaps."""
    n = 3000
    A, kw = workloads.shifted_goe(n, 17, 0.97)
    for method in ('maximal_stability', 'minimal_difference'):
        kwm = dict(kw, permutation=method)
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kwm)
        assert not np.array_equal(np.asarray(dec.p), np.arange(n))
        _check_identity_and_bounds(A, dec, kwm)


# ---------------------------------------------------------------------------------------------
# GMW family (GMW_SE.py): modified LDL^T without pivoting, d_k from |c|_inf of the current column
# ---------------------------------------------------------------------------------------------
def test_gmw_family_against_reference_goldens(matrix):
    """GMW_81 / GMW_T1 / GMW_T2 on the device against the 96 golden cases of the unmodified reference.  The kernels
    mirror numpy operation by operation, so B agrees to the last bit except where max-reductions meet exact ties."""
    z = load_golden('gmw.npz')
    manifest = json.loads(str(z['manifest']))
    aps = matrix.approximation.positive_semidefinite
    n_exact = 0
    for m in manifest:
        A = z[m['matrix'] + '/A']
        kw = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m['kwargs'].items()}
        A0 = A.copy()
        B = getattr(aps, m['fn'])(A, **kw)
        assert np.array_equal(A, A0)                                   # A untouched (overwrite_A=False)
        ref = z[m['key'] + '/B']
        np.testing.assert_allclose(B, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max(), err_msg=m['key'])
        n_exact += int(np.array_equal(B, ref))
    assert n_exact >= 0.9 * len(manifest), n_exact


def test_gmw_family_against_oracle_and_properties(matrix):
    from oracle import gmw_oracle
    aps = matrix.approximation.positive_semidefinite
    rng = np.random.default_rng(11)
    for n in (1, 5, 257, 700):
        g = rng.standard_normal((n, n))
        A = (g + g.T) / 2
        for fn in ('GMW_81', 'GMW_T1', 'GMW_T2'):
            B = getattr(aps, fn)(A)
            ref = gmw_oracle.approximate(A, fn)
            np.testing.assert_allclose(B, ref, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(ref).max()))
            off = ~np.eye(n, dtype=bool)
            assert np.array_equal(B[off], A[off])                      # only the diagonal is modified
            assert np.all(B.diagonal() >= A.diagonal())
            if n > 1:
                assert np.linalg.eigvalsh(B).min() > 0                 # positive definite
    S = workloads.lowrank_plus_diag(300, 2)                           # already positive definite: unchanged by T1 / T2
    for fn in ('GMW_T1', 'GMW_T2'):
        np.testing.assert_allclose(getattr(aps, fn)(S), S, rtol=0, atol=1e-12)
    with pytest.raises(matrix.errors.MatrixNotSquareError):
        aps.GMW_81(np.ones((3, 4)))


def test_se_family_against_reference_goldens_and_oracle(matrix):
    """SE_90 / SE_99 / SE_T1 (SE_T2 = SE_99) on the device against the 120 golden cases of the reference
    (tests/golden/se.npz, see generate_se_goldens.py for how the reference was run) and, at sizes the goldens do not
    reach, against the oracle.  Tolerance 1e-10 relative to max |B|: the column 1-norm is a tree sum here and numpy's
    pairwise sum there, the 2 x 2 eigenvalues are a closed form here and LAPACK there."""
    from oracle import gmw_oracle
    z = load_golden('se.npz')
    manifest = json.loads(str(z['manifest']))
    aps = matrix.approximation.positive_semidefinite
    for m in manifest:
        A = z[m['matrix'] + '/A']
        kw = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m['kwargs'].items()}
        A0 = A.copy()
        B = getattr(aps, m['fn'])(A, **kw)
        assert np.array_equal(A, A0)
        ref = z[m['key'] + '/B']
        np.testing.assert_allclose(B, ref, rtol=TOL, atol=TOL * np.abs(ref).max(), err_msg=m['key'])
    rng = np.random.default_rng(12)
    for n in (5, 257, 700):
        g = rng.standard_normal((n, n))
        A = (g + g.T) / 2
        for fn in ('SE_90', 'SE_99', 'SE_T1', 'SE_T2'):
            B = getattr(aps, fn)(A)
            ref = gmw_oracle.approximate(A, fn)
            np.testing.assert_allclose(B, ref, rtol=TOL, atol=TOL * max(1.0, np.abs(ref).max()))
            off = ~np.eye(n, dtype=bool)
            assert np.array_equal(B[off], A[off])                      # only the diagonal is modified
            assert np.all(B.diagonal() >= A.diagonal())
            assert np.linalg.eigvalsh(B).min() > 0                     # positive definite
    S = workloads.lowrank_plus_diag(300, 2)                           # already positive definite: unchanged
    for fn in ('SE_90', 'SE_99', 'SE_T1'):
        np.testing.assert_allclose(getattr(aps, fn)(S), S, rtol=0, atol=1e-12)


# ---------------------------------------------------------------------------------------------
# decompose, solve, multiply, permute
# ---------------------------------------------------------------------------------------------
def test_decompose_solve_multiply_against_reference_goldens(matrix):
    z = load_golden('decompose_solve_permute.npz')
    manifest = json.loads(str(z['manifest']))
    A = z['spd10/A']
    for m in manifest:
        key = m['key']
        if not key.startswith('spd10/'):
            continue
        perm = m['permutation']
        perm = np.asarray(perm, dtype=np.int64) if isinstance(perm, list) else perm
        A_copy = A.copy()
        dec = matrix.decompose(A, permutation=perm, return_type=m['return_type'])
        assert np.array_equal(A, A_copy)
        np.testing.assert_array_equal(np.asarray(dec.p), z[key + '/p'])
        if m['return_type'] == 'LL':
            np.testing.assert_allclose(dec.L, z[key + '/L'], rtol=TOL, atol=1e-13)
        elif m['return_type'] == 'LDL':
            np.testing.assert_allclose(dec.L, z[key + '/L'], rtol=TOL, atol=1e-13)
            np.testing.assert_allclose(dec.d, z[key + '/d'], rtol=TOL)
        else:
            np.testing.assert_allclose(dec.LD, z[key + '/LD'], rtol=TOL, atol=1e-13)
        np.testing.assert_allclose(dec.solve(z['spd10/b1']), z[key + '/x1'], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dec.solve(z['spd10/b2']), z[key + '/x2'], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dec.matrix_right_side_multiplication(z['spd10/b2']), z[key + '/Ab2'], rtol=1e-11,
                                   atol=1e-11)
        assert matrix.util.is_almost_equal(dec.composed_matrix, A, atol=1e-6)          # test_decompose.py:40
    A = workloads.lowrank_plus_diag(128, 5)
    dec = matrix.decompose(A, permutation='decreasing_diagonal_values', return_type='LDL')
    np.testing.assert_array_equal(np.asarray(dec.p), z['f3_128/p'])
    np.testing.assert_allclose(dec.L, z['f3_128/L'], rtol=TOL, atol=1e-13)
    np.testing.assert_allclose(dec.d, z['f3_128/d'], rtol=TOL)
    np.testing.assert_allclose(dec.solve(z['f3_128/b']), z['f3_128/x'], rtol=TOL, atol=1e-12)
    np.testing.assert_allclose(dec.solve(z['f3_128/b'][:, 0]), z['f3_128/x0'], rtol=TOL, atol=1e-12)


def test_factor_stays_on_device_until_read(matrix, orc):
    """decompose / approximate keep the n x n factor in HBM: solve and multiply work on it there, .L downloads
    it on first access -- also after a solve has mirrored it into the upper triangle on the device."""
    n = 700
    S = workloads.lowrank_plus_diag(n, 8)
    b = np.random.default_rng(1).standard_normal(n)
    for rt in ('LDL', 'LL', 'LDL_compressed'):
        dec = matrix.decompose(S, permutation='decreasing_diagonal_values', return_type=rt)
        assert dec.is_device_resident_only and dec.n == n
        x = dec.solve(b)
        assert dec.is_device_resident_only                      # solve did not need the host copy
        np.testing.assert_allclose(S @ x, b, rtol=0, atol=1e-9)
        ref = orc.decompose(S, 'decreasing_diagonal_values', rt)
        M = dec.LD if rt == 'LDL_compressed' else dec.L         # first access: download (after the solve)
        assert not dec.is_device_resident_only
        assert np.all(np.triu(M, 1) == 0)
        if rt == 'LDL_compressed':
            np.testing.assert_allclose(np.tril(M, -1), np.tril(ref['L'], -1), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(M.diagonal(), ref['d'], rtol=1e-10)
        else:
            np.testing.assert_allclose(M, ref['L'], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(dec.solve(b), x, rtol=1e-12, atol=1e-14)   # still solves after the download
        import copy
        dec2 = matrix.decompose(S, permutation='decreasing_diagonal_values', return_type=rt)
        dec3 = copy.deepcopy(dec2)                               # copying materialises
        assert not dec2.is_device_resident_only and dec3.is_equal(dec)
    A, kw = workloads.shifted_goe(300, 4, 1.05)
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    assert dec.is_device_resident_only
    ref = orc.approximate(A, **kw)
    np.testing.assert_allclose(np.asarray(dec.d, dtype=float), ref['d'].astype(float), rtol=1e-10)
    assert np.all(np.abs(dec.L - ref['L']) <= 1e-10 * np.maximum(1.0, np.abs(ref['L'])))
    # the finite check of large inputs happens on the device (small ones are scanned on the host)
    from matrix_decomposition_b200 import calculate
    S2 = workloads.lowrank_plus_diag(1100, 9)
    assert S2.size > calculate._HOST_FINITE_SCAN_MAX
    with pytest.raises(matrix.errors.MatrixNotFiniteError):
        Sbad = S2.copy()
        Sbad[5, 7] = Sbad[7, 5] = np.inf
        matrix.decompose(Sbad)
    Sbad = S2.copy()
    Sbad[3, 3] = np.nan
    with pytest.raises(matrix.errors.MatrixNotFiniteError):
        matrix.decompose(Sbad, permutation='none')


@pytest.mark.parametrize('rt', ['LDL', 'LL'])
def test_one_shot_streamed_output_equals_resident_factor(matrix, rt):
    """From n = 4096 on the one-shot C-ABI calls upload only the upper triangle of A and stream finished rows of L
    to the host while the factorization is still running; the result must equal the factor that stayed on the
    device (downloaded afterwards), bit for bit, and satisfy the oracle-free identity of Reimer.py:947-957."""
    from matrix_decomposition_b200 import _native
    ctx = _native.context()
    lib = ctx.lib
    n = 4300
    A, kw = workloads.shifted_goe(n, 77, 1.05)
    p = np.argsort(-A.diagonal()).astype(np.int64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    L = ctx.pinned_array((n, n))          # streaming needs pinned memory (pageable L takes the one-copy path)
    L[:] = np.nan
    d, om, de = np.empty(n), np.empty(n), np.empty(n)
    cl = np.zeros(n, dtype=np.uint8)
    h = ctypes.c_void_p()
    P = _native.ptr
    ctx.check(lib.mdb_approximate_f64(ctx.handle, n, P(A), n, P(p), ctypes.byref(bounds), _native.TYPE_CODES[rt], P(L), n,
                                      P(d), P(om), P(de), P(cl), _native.HOST, ctypes.byref(h)), 'one-shot')
    resident = _native.Factor(ctx, h).download(n, _native.TYPE_CODES[rt])
    assert np.array_equal(L, resident)
    Lp = np.full((n, n), np.nan)          # the same call with a pageable L
    ctx.check(lib.mdb_approximate_f64(ctx.handle, n, P(A), n, P(p), ctypes.byref(bounds), _native.TYPE_CODES[rt], P(Lp), n,
                                      P(d), P(om), P(de), P(cl), _native.HOST, None), 'one-shot pageable')
    assert np.array_equal(Lp, resident)
    Lpin = L
    L = np.array(Lpin)                     # off the pinned buffer, which is released right away
    _native.free_pinned(Lpin)
    assert np.all(np.triu(L, 1) == 0)
    # identity: P^T L D L^T P = A o Omega + diag(delta)  on a probe vector
    v = np.random.default_rng(3).standard_normal(n)
    Lu = L if rt == 'LDL' else L / np.sqrt(d)[None, :]
    lhs = Lu @ (d * (Lu.T @ v))
    Ap = A[p[:, None], p[None, :]]
    om_pos, de_pos = om[p], de[p]
    later = np.maximum(np.arange(n)[:, None], np.arange(n)[None, :])
    M = Ap * om_pos[later]
    M[np.diag_indices(n)] = Ap.diagonal() + de_pos
    rhs = M @ v
    assert np.max(np.abs(lhs - rhs)) <= 1e-11 * np.max(np.abs(Ap @ v))
    assert 0.2 * n < cl.sum() < 0.7 * n


def test_config2_size_decompose_and_solve_properties(matrix):
    """BASELINE config 2 at full size (SPD, n = 16384, LDL, solve with 1 and 64 right-hand sides), checked through
    size-independent properties: A x = b, the composed matrix applied to probes equals A applied to them, d > 0,
    L unit lower triangular; the factor never leaves the device except for the final look at L."""
    n = 16384
    A = workloads.lowrank_plus_diag(n, 12)
    dec = matrix.decompose(A, permutation='decreasing_diagonal_values', return_type='LDL')
    assert dec.is_device_resident_only
    rng = np.random.default_rng(2)
    b = rng.standard_normal(n)
    x = dec.solve(b)                                            # persistent triangular sweeps (1 rhs)
    assert np.max(np.abs(A @ x - b)) <= 1e-10 * np.max(np.abs(b)) * n ** 0.5
    Bk = rng.standard_normal((n, 64))
    Xk = dec.solve(Bk)                                          # blocked DMMA sweeps
    assert np.max(np.abs(A @ Xk - Bk)) <= 1e-10 * np.max(np.abs(Bk)) * n ** 0.5
    np.testing.assert_allclose(dec.solve(Bk[:, 0]), Xk[:, 0], rtol=1e-9, atol=1e-12)   # both paths agree
    V = rng.standard_normal((n, 3))
    AV = dec.matrix_right_side_multiplication(V)                # (P^T L D L^T P) V on the device
    assert np.max(np.abs(AV - A @ V)) <= 1e-11 * np.max(np.abs(A @ V))
    d = np.asarray(dec.d, dtype=float)
    assert np.all(d > 0) and dec.is_positive_definite()
    assert sorted(np.asarray(dec.p).tolist()) == list(range(n))
    assert np.all(np.diff(A.diagonal()[np.asarray(dec.p)]) <= 0)     # 'decreasing_diagonal_values'
    L = dec.L
    assert np.all(L.diagonal() == 1) and np.all(np.triu(L[:2048, :2048], 1) == 0) and L[5, n - 1] == 0


def test_decompose_failure_index(matrix):
    z = load_golden('decompose_solve_permute.npz')
    with pytest.raises(matrix.errors.NoDecompositionPossibleWithProblematicSubdecompositionError) as ei:
        matrix.decompose(z['indef10/A'], permutation='none')
    assert ei.value.problematic_leading_principal_submatrix_index == int(z['indef10/bad_index'])
    assert not matrix.is_positive_definite(z['indef10/A'])
    assert matrix.is_positive_definite(z['spd10/A']) and matrix.is_invertible(z['spd10/A'])


@pytest.mark.parametrize('n,nrhs', [(300, 1), (300, 2), (300, 9), (1000, 1), (1000, 130), (2500, 1), (2500, 24)])
def test_solve_and_multiply_against_oracle(matrix, orc, n, nrhs):
    A = workloads.lowrank_plus_diag(n, n)
    for rt in ('LDL', 'LL'):
        dec = matrix.decompose(A, permutation='decreasing_diagonal_values', return_type=rt)
        rng = np.random.default_rng(nrhs)
        b = rng.standard_normal(n) if nrhs == 1 else rng.standard_normal((n, nrhs))
        x = dec.solve(b)
        d = dec.d if rt == 'LDL' else None
        x_ref = orc.solve(dec.L, d, np.asarray(dec.p), b)
        assert x.shape == b.shape
        assert np.max(np.abs(x - x_ref)) <= TOL * np.max(np.abs(x_ref))
        assert np.max(np.abs(A @ x - b)) <= 1e-9 * np.max(np.abs(b))                    # test_solve.py:31
        y = dec.matrix_right_side_multiplication(b)
        assert np.max(np.abs(y - A @ b)) <= 1e-10 * np.max(np.abs(A @ b))               # test_multiply.py:30
        # an uploaded factor (no resident handle) gives the same answer
        fresh = dec.copy()
        assert np.array_equal(fresh.solve(b), x)


@pytest.mark.parametrize('n', [1, 100, 128, 700, 8300])
def test_composed_matrix_on_device(matrix, n):
    """DecompositionBase.composed_matrix (decompositions.py:826-830 LDL, :1202-1205 LL) is formed on the device from
    the resident factor (mdb_factor_composed): against numpy's P^T (L D L^T) P for every decomposition type, with and
    without a permutation; n = 8300 takes the two-pass permutation and several waves of the tile kernel.  Symmetric bit
    for bit (the upper triangle is the mirror image of the lower one)."""
    rng = np.random.default_rng(n)
    L = np.tril(rng.standard_normal((n, n)) / np.sqrt(n), -1) + np.eye(n)
    d = rng.uniform(0.5, 2.0, n)
    for p in (None, rng.permutation(n)):
        want = (L * d[None, :]) @ L.T
        if p is not None:
            full = np.empty_like(want)
            full[np.ix_(p, p)] = want
            want = full
        ldl = matrix.decompositions.LDL_Decomposition(L, d, p=p)
        decs = [ldl] if n > 1000 else [ldl, ldl.as_type('LDL_compressed'), ldl.as_type('LL')]
        for dec in decs:
            B = dec.composed_matrix
            assert B.shape == (n, n) and np.array_equal(B, B.T)
            assert np.max(np.abs(B - want)) <= 1e-12 * np.max(np.abs(want)), (n, dec.type_str)


def test_solve_zero_and_integer_rhs_and_singular(matrix):
    A = workloads.reference_fixture_spd(10)
    dec = matrix.decompose(A, return_type='LDL')
    assert np.all(dec.solve(np.zeros(10)) == 0)
    x = dec.solve(np.arange(10))
    np.testing.assert_allclose(A @ x, np.arange(10), rtol=1e-9, atol=1e-9)
    sing = matrix.decompositions.LDL_Decomposition(np.eye(4), np.array([1.0, 0.0, 1.0, 1.0]))
    with pytest.raises(matrix.errors.DecompositionSingularError):
        sing.solve(np.ones(4))
    both = dec.inverse_matrix_both_sides_multiplication(np.arange(10.0))
    np.testing.assert_allclose(both, np.arange(10.0) @ np.linalg.solve(A, np.arange(10.0)), rtol=1e-9)
    both = dec.matrix_both_sides_multiplication(np.arange(10.0), np.ones(10))
    np.testing.assert_allclose(both, np.ones(10) @ A @ np.arange(10.0), rtol=1e-10)


def test_permute_symmetric_bit_exact(matrix):
    z = load_golden('decompose_solve_permute.npz')
    np.testing.assert_array_equal(matrix.permute.symmetric(z['permute10/A'], z['permute10/p']), z['permute10/B'])
    for n in (1, 33, 1024, 1500):
        A = workloads.goe(n, n)
        p = np.random.default_rng(n).permutation(n)
        B = matrix.permute.symmetric(A, p)
        assert np.array_equal(B, A[p[:, None], p[None, :]])                              # test_permute.py:60
        assert np.array_equal(matrix.permute.symmetric(B, matrix.permute.invert_permutation_vector(p)), A)
    # above 8192 rows the permutation runs as a scatter with the inverse permutation (k_scatter_sym): sizes that do and
    # do not fill the last chunk of 8192 source columns, a random permutation, the identity and a reversal
    for n in (8193, 9000, 16384):
        rng = np.random.default_rng(n)
        A = rng.standard_normal((n, n))
        for p in (rng.permutation(n), np.arange(n), np.arange(n)[::-1].copy()):
            B = matrix.permute.symmetric(A, p)
            assert np.array_equal(B, A[p[:, None], p[None, :]])


@pytest.mark.parametrize('shift,rt,n,zero_copy', [(1.05, 'LDL', 4300, True), (0.9, 'LDL', 4300, True),
                                                  (0.9, 'LDL_compressed', 4300, True), (0.9, 'LDL', 4301, True),
                                                  (0.9, 'LDL', 4300, False), (0.9, 'LDL', 16500, True)])
def test_one_shot_early_shipment_of_late_rows(matrix, shift, rt, n, zero_copy, monkeypatch):
    """Large one-shot calls ship the columns left of the last quarter of the rows before those rows are final (unscaled
    values, equal to L unless omega != 1 or an entry is lost to the threshold) and send the rows whose early part turned
    out different again (mdb_factor_run, k_finalize's `changed` flags; k_resend_rows writes them into the pinned matrix
    from the device, n = 4301: rows of the host matrix not 16-byte aligned).  Forced on at n = 4300 (MDB200_TAIL_MIN),
    as shipped at n = 16500: with
    s = 1.05 every omega is 1 (nothing is re-sent), with s = 0.9 rows are shrunk (omega < 1) and must be re-sent; the
    streamed L must equal the factor that stayed on the device, bit for bit."""
    from matrix_decomposition_b200 import _native
    if n < 16384:
        monkeypatch.setenv('MDB200_TAIL_MIN', '1024')      # (n = 16500 runs it as shipped: on from n = 16384)
    if not zero_copy:    # the re-send through copies issued by the host at the end (host matrix not device-addressable)
        monkeypatch.setenv('MDB200_NO_ZEROCOPY', '1')
    ctx = _native.context()
    lib = ctx.lib
    A, kw = workloads.shifted_goe(n, 78, shift)
    p = np.argsort(-A.diagonal()).astype(np.int64)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    L = ctx.pinned_array((n, n))
    L[:] = np.nan
    d, om, de = np.empty(n), np.empty(n), np.empty(n)
    cl = np.zeros(n, dtype=np.uint8)
    h = ctypes.c_void_p()
    P = _native.ptr
    ctx.check(lib.mdb_approximate_f64(ctx.handle, n, P(A), n, P(p), ctypes.byref(bounds), _native.TYPE_CODES[rt], P(L), n,
                                      P(d), P(om), P(de), P(cl), _native.HOST, ctypes.byref(h)), 'one-shot')
    resident = _native.Factor(ctx, h).download(n, _native.TYPE_CODES[rt])
    same = np.array_equal(L, resident)
    shrunk = int(np.count_nonzero(om[p][n - n // 4:] != 1.0))
    _native.free_pinned(L)
    assert same
    assert (shrunk == 0) if shift > 1 else (shrunk > 10)      # the second case really exercises the re-send


# ---------------------------------------------------------------------------------------------
# the tile kernel, the device generator, and the full-size identity
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K,lower,beta1,neg', [(128, 64, 16, 0, 0, 0), (1, 1, 16, 0, 1, 0), (257, 191, 128, 0, 1, 1),
                                                   (1000, 1000, 128, 1, 1, 0), (333, 777, 48, 1, 1, 0)])
def test_dmma_tile_kernel(M, N, K, lower, beta1, neg):
    from matrix_decomposition_b200 import _native
    ctx = _native.context()
    rng = np.random.default_rng(M + N + K)
    A, B, C0 = rng.standard_normal((M, K)), rng.standard_normal((N, K)), rng.standard_normal((M, N))
    C = C0.copy()
    ms = ctypes.c_double()
    rg, cg = (64, 0) if lower else (0, 0)
    P = _native.ptr
    ctx.check(ctx.lib.mdb_debug_gemm_tn(ctx.handle, M, N, K, P(A), K, P(B), K, P(C), N, lower, beta1, neg, rg, cg, 1,
                                        ctypes.byref(ms)), 'mdb_debug_gemm_tn')
    prod = A @ B.T
    want = (C0 if beta1 else 0) + (-prod if neg else prod)
    if lower:
        want = np.where((rg + np.arange(M))[:, None] > (cg + np.arange(N))[None, :], want, C0)
    assert np.max(np.abs(C - want)) <= 1e-12 * K


def test_device_generator_and_identity_at_scale():
    """Family F2 generated on the device at n = 8192: symmetric, right statistics; the factorization of it
    satisfies the oracle-free identity L D L^T = A o Omega + diag(delta) (Reimer.py:947-957) to rounding."""
    from matrix_decomposition_b200 import _native
    ctx = _native.context()
    lib = ctx.lib
    n = 8192
    mu = 1.05 * math.sqrt(2 * n)
    dA = ctx.malloc(n * n * 8)
    ctx.check(lib.mdb_generate_f2(ctx.handle, n, 11, mu, ctypes.c_void_p(dA), n), 'gen')
    A = np.empty((n, n))
    ctx.check(lib.mdb_memcpy(ctx.handle, _native.ptr(A), ctypes.c_void_p(dA), n * n * 8, _native.HOST, _native.DEVICE), 'copy')
    ctx.free(dA)
    assert np.array_equal(A, A.T)
    off = A[np.triu_indices(n, 1)]
    assert abs(off.mean()) < 2e-3 and abs(off.var() - 0.5) < 5e-3
    assert abs((A.diagonal() - mu).var() - 1.0) < 0.1
    diag = np.empty(n)
    ctx.check(lib.mdb_generate_diag_f2(ctx.handle, n, 11, mu, _native.ptr(diag)), 'diag')
    assert np.array_equal(diag, A.diagonal())
    p = np.argsort(-diag).astype(np.int64)
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'create')
    ctx.check(lib.mdb_factor_generate_f2(f, 11, mu, _native.ptr(p)), 'gen2')
    kw = workloads.f2_bounds(n)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
    res = ctypes.c_double()
    ctx.check(lib.mdb_factor_identity_residual(f, 3, 5, ctypes.byref(res)), 'identity')
    assert res.value <= 1e-12
    # same matrix through the host API gives the same decisions as the device-resident path
    d = np.empty(n)
    cl = np.zeros(n, dtype=np.uint8)
    ctx.check(lib.mdb_factor_get(f, _native.TYPE_LDL, None, 0, _native.HOST, _native.ptr(d), None, None, _native.ptr(cl),
                                 None), 'get')
    lib.mdb_factor_destroy(f)
    import matrix_decomposition_b200 as matrix
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    assert np.array_equal(dec.clamped, cl.astype(bool)) and np.array_equal(np.asarray(dec.d, dtype=float), d)
    assert 0.2 * n < cl.sum() < 0.6 * n


def test_buffer_cache_reuse_and_trim(matrix):
    """The context keeps the working matrix / staging buffer of a large call for the next one (mdb_ctx_trim,
    include/mdb200.h): results must not depend on what the reused buffers held before."""
    from matrix_decomposition_b200 import _native
    n = 6100                                    # > 256 MB working matrix, not a multiple of the block size
    A1, kw = workloads.shifted_goe(n, 31, 1.05)
    A2, _ = workloads.shifted_goe(n, 32, 0.95)
    aps = matrix.approximation.positive_semidefinite
    matrix.release_device_memory()
    fresh = aps.decomposition(A1, **kw)
    L_fresh, d_fresh = np.array(fresh.L), np.asarray(fresh.d, dtype=float).copy()
    del fresh
    other = aps.decomposition(A2, **kw)         # reuses the buffers, leaves different contents behind
    _ = other.L
    del other
    again = aps.decomposition(A1, **kw)
    assert np.array_equal(np.asarray(again.d, dtype=float), d_fresh) and np.array_equal(again.L, L_fresh)
    del again
    matrix.release_device_memory()
    _native.context().trim()                    # idempotent
    once_more = aps.decomposition(A1, **kw)
    assert np.array_equal(once_more.L, L_fresh)


def test_config3_full_size_identity_and_bounds():
    """BASELINE config 3 (n = 65536, 34 GB, every outer-block width 8 / 4 / 2 / 1 in play): the oracle cannot run
    here, so the result is checked through the size-independent identity L D L^T = A o Omega + diag(delta)
    (Reimer.py:947-957) on random probes and through the bounds every d has to respect (Reimer.py:64-67)."""
    from matrix_decomposition_b200 import _native
    ctx = _native.context()
    lib = ctx.lib
    n = 65536
    mu = 1.05 * math.sqrt(2 * n)
    diag = np.empty(n)
    ctx.check(lib.mdb_generate_diag_f2(ctx.handle, n, 3, mu, _native.ptr(diag)), 'diag')
    p = np.argsort(-diag).astype(np.int64)
    f = ctypes.c_void_p()
    ctx.check(lib.mdb_factor_create(ctx.handle, n, 1, ctypes.byref(f)), 'create')
    try:
        ctx.check(lib.mdb_factor_generate_f2(f, 3, mu, _native.ptr(p)), 'gen')
        kw = workloads.f2_bounds(n)
        bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                           float(np.finfo(float).eps), None, kw['max_diag_B'])
        ctx.check(lib.mdb_factor_run(f, ctypes.byref(bounds)), 'run')
        res = ctypes.c_double()
        ctx.check(lib.mdb_factor_identity_residual(f, 2, 9, ctypes.byref(res)), 'identity')
        assert res.value <= 1e-12
        d, om, de = np.empty(n), np.empty(n), np.empty(n)
        cl = np.zeros(n, dtype=np.uint8)
        ctx.check(lib.mdb_factor_get(f, _native.TYPE_LDL, None, 0, _native.HOST, _native.ptr(d), _native.ptr(om),
                                     _native.ptr(de), _native.ptr(cl), None), 'get')
    finally:
        lib.mdb_factor_destroy(f)
    lo = max(kw['min_diag_D'], kw['min_abs_value_D'])
    assert np.all(np.isfinite(d)) and d.min() >= lo and d.max() <= kw['max_diag_D']
    assert np.all((om >= 0) & (om <= 1)) and np.all(np.isfinite(de))
    cl = cl.astype(bool)
    assert np.all(om[p][~cl] == 1.0)                  # unclamped <=> the rule stayed in its first branch
    assert 0.2 * n < cl.sum() < 0.6 * n


@pytest.mark.parametrize('n,fixed', [(700, 1), (1500, 4), (1100, 8)])
def test_distributed_steps_single_rank_equal_single_gpu_path(matrix, orc, n, fixed):
    """The multi-GPU step functions with world = 1 (no collective) reproduce the single-GPU path bit for bit,
    with and without outer-block aggregation."""
    import torch
    from matrix_decomposition_b200 import distributed as md, _native
    ctx = _native.context()
    A, kw = workloads.shifted_goe(n, 3, 1.05)

    class LocalComm:
        def broadcast(self, tensor, src, stream):
            class H:
                def wait(self, s):
                    pass
            return H()

    try:
        ctx.set_blocking(fixed=fixed)
        dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        p = np.asarray(dec.p)
        Ap = A[p[:, None], p[None, :]]
        main, bulk = md._streams(torch.cuda.current_device())
        steps = md.CudaSteps(n, 0, 1)
        assert steps.outer_width(0) == fixed
        steps.set_matrix(Ap, main)
        bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                           float(np.finfo(float).eps), None, kw['max_diag_B'])
        steps.set_bounds(bounds, p, main)
        md.factorize(n, 0, 1, steps, LocalComm(), main, bulk, lookahead=True)
        torch.cuda.synchronize()
        d, omega, delta, clamped = steps.finalize(main)
        L = steps.S.cpu().numpy()[:, :n]
        steps.close()
    finally:
        ctx.set_blocking()
    assert np.array_equal(clamped, dec.clamped)
    assert np.array_equal(d, np.asarray(dec.d, dtype=float))
    assert np.array_equal(L, dec.L)


@pytest.mark.parametrize('n', [8192 + 37, 16384 + 37])
def test_decisions_match_oracle_with_default_blocking_at_scale(matrix, orc, n):
    """Decision parity above the golden range, with the DEFAULT outer-block schedule (aggregation of 2 / 4 panels
    starts at 8192 / 16384 remaining rows): permutation and clamped set identical to the oracle, d / L element-wise
    within 1e-10, B = L D L^T within 1e-12 (both composed on the device from the respective factors).  The oracle
    needs ~6 s / ~60 s on the box's host cores."""
    A, kw = workloads.shifted_goe(n, 9000 + n, 1.05)
    orc.set_threads(len(os.sched_getaffinity(0)))
    ref = orc.approximate(A, **kw)
    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    np.testing.assert_array_equal(np.asarray(dec.p), ref['p'])
    np.testing.assert_array_equal(dec.clamped, ref['clamped'])
    assert 0.2 * n < ref['clamped'].sum() < 0.8 * n
    d_ref = ref['d'].astype(np.float64)
    np.testing.assert_allclose(np.asarray(dec.d, dtype=np.float64), d_ref, rtol=TOL)
    np.testing.assert_allclose(np.asarray(dec.omega, dtype=np.float64), ref['omega'].astype(np.float64), rtol=TOL, atol=1e-14)
    B = dec.composed_matrix
    L = dec.L
    assert np.all(np.abs(L - ref['L']) <= TOL * np.maximum(1.0, np.abs(ref['L'])))
    del L
    B_ref = matrix.decompositions.LDL_Decomposition(ref['L'], d_ref, p=ref['p']).composed_matrix
    assert np.linalg.norm(B - B_ref) <= TOL_B * np.linalg.norm(B_ref)


def test_batched_python_entry_equals_single_calls(matrix):
    """decomposition_batched (config 5's entry point) returns, per matrix, what decomposition returns."""
    psd = matrix.approximation.positive_semidefinite
    n, batch = 200, 5
    rng = np.random.default_rng(11)
    G = rng.standard_normal((batch, n, n))
    mu = 1.05 * math.sqrt(2 * n)
    As = (G + G.transpose(0, 2, 1)) / 2 + mu * np.eye(n)[None]
    kw = workloads.f2_bounds(n)
    res = psd.decomposition_batched(As, **kw)
    assert len(res) == batch and res.L.shape == (batch, n, n)
    for b in range(batch):
        one = psd.decomposition(As[b], **kw)
        dec = res[b]
        assert isinstance(dec, matrix.decompositions.LDL_Decomposition)
        np.testing.assert_array_equal(dec.p, one.p)
        np.testing.assert_array_equal(dec.clamped, one.clamped)
        np.testing.assert_array_equal(np.asarray(dec.d, dtype=float), np.asarray(one.d, dtype=float))
        np.testing.assert_array_equal(dec.L, one.L)
        np.testing.assert_array_equal(np.asarray(dec.omega, dtype=float), np.asarray(one.omega, dtype=float))
        np.testing.assert_array_equal(np.asarray(dec.delta, dtype=float), np.asarray(one.delta, dtype=float))
    # explicit permutation (one vector for all), vector bound, shard helper
    perm = np.random.default_rng(1).permutation(n)
    res2 = psd.decomposition_batched(As[:2], min_diag_D=0.5, max_diag_B=np.full(n, 3 * mu), permutation=perm)
    one2 = psd.decomposition(As[1], min_diag_D=0.5, max_diag_B=np.full(n, 3 * mu), permutation=perm)
    np.testing.assert_array_equal(res2[1].L, one2.L)
    assert [psd.shard_for_rank(10, r, 4) for r in range(4)] == [slice(0, 3), slice(3, 6), slice(6, 8), slice(8, 10)]
    with pytest.raises(matrix.errors.MatrixNotSquareError):
        psd.decomposition_batched(np.zeros((2, 3, 4)))
    # dynamic pivoting falls back to per-matrix calls
    res3 = psd.decomposition_batched(As[:2, :60, :60], min_diag_D=0.5)
    one3 = psd.decomposition(As[1, :60, :60], min_diag_D=0.5)
    np.testing.assert_array_equal(res3[1].p, one3.p)
    np.testing.assert_array_equal(res3[1].L, one3.L)


@pytest.mark.parametrize('n', [700, 1500, 2600])
def test_distributed_factor_consumers_single_rank(matrix, n):
    """The consumers of the block-cyclic factor (mdb_dist_solve_*, mdb_dist_place_columns, mdb_factor_upload_device)
    with world = 1: solve on the slab against the single-GPU solve, gather bit-identical to the single-GPU L."""
    import torch
    from matrix_decomposition_b200 import distributed as md, _native
    A, kw = workloads.shifted_goe(n, 3, 1.05)

    class LocalComm:
        def broadcast(self, tensor, src, stream):
            class H:
                def wait(self, s):
                    pass
            return H()

    dec = matrix.approximation.positive_semidefinite.decomposition(A, **kw)
    p = np.asarray(dec.p)
    Ap = A[p[:, None], p[None, :]]
    main, bulk = md._streams(torch.cuda.current_device())
    steps = md.CudaSteps(n, 0, 1)
    steps.set_matrix(Ap, main)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, kw['min_diag_D'], kw['max_diag_D'], kw['min_abs_value_D'],
                                       float(np.finfo(float).eps), None, kw['max_diag_B'])
    steps.set_bounds(bounds, p, main)
    md.factorize(n, 0, 1, steps, LocalComm(), main, bulk, lookahead=True)
    torch.cuda.synchronize()
    d, omega, delta, clamped = steps.finalize(main)
    fac = md.DistributedFactor(steps, d, p, omega, delta, clamped)
    rng = np.random.default_rng(2)
    for shape in ((n,), (n, 3), (n, 11)):
        b = rng.standard_normal(shape)
        x = fac.solve(b)
        x_ref = dec.solve(b)
        assert x.shape == b.shape
        assert np.max(np.abs(x - x_ref)) <= TOL * np.max(np.abs(x_ref))
    g = fac.gather(0)
    np.testing.assert_array_equal(g.L, dec.L)
    np.testing.assert_array_equal(np.asarray(g.omega, dtype=float), np.asarray(dec.omega, dtype=float))
    np.testing.assert_array_equal(np.asarray(g.delta, dtype=float), np.asarray(dec.delta, dtype=float))
    b = rng.standard_normal(n)
    np.testing.assert_array_equal(g.solve(b), dec.solve(b))
    fac.close()


def test_two_gpu_block_cyclic_equals_single_gpu(tmp_path):
    """Needs two GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2`): tools/dist_check.py under torchrun --
    the 1-D block-cyclic factorization with the NCCL panel broadcast must reproduce the single-GPU factor bit for bit
    (outer blocks of 1, 2 and 4 panels, with and without lookahead) and satisfy the identity at a size with many outer
    blocks."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, DIST_N='4096')
    env.pop('MDB200_DEVICE', None)
    port = 29600
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', str(port),
                        os.path.join(REPO, 'tools', 'dist_check.py')], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.load(open(os.path.join(REPO, 'gpurun_out', 'dist_check_2gpu.json')))
    cases = [v for k, v in out.items() if k.startswith('n')]
    assert len(cases) == 6 and all(c['bitwise_equal_single_gpu'] and c['identity'] <= 1e-12 for c in cases)
    # the consumers of the distributed factor: gather bit-identical, solve on the slabs within 1e-10
    assert all(c['gather_bitwise_equal'] and c['gathered_solve_equal'] for c in cases)
    assert all(c['dist_solve_rel_err'] <= TOL and c['dist_solve_1rhs_rel_err'] <= TOL for c in cases)
    big = out['bench_n4096_gpus2']
    assert big['identity'] <= 1e-12
    assert big['dist_solve']['residual_vs_original_matrix'] <= 1e-10
    assert big['panel_exchange'] == 'NCCL broadcast'
