"""The blocked right-looking restatement (oracle/blocked_model.py, the numpy twin of the CUDA
kernels) and the FP64 scalar rule header (csrc/md_rule.h) against the oracle and the goldens."""

import numpy as np
import pytest

import workloads
from oracle import blocked_model as bm
from oracle import cpu_oracle as orc
from conftest import load_golden, ld_join


def test_fp64_rule_against_reference_vectors():
    """csrc/md_rule.h (the header the kernels compile) on the reference's own _minimal_change outputs."""
    z = load_golden('scalar_rule.npz')
    want = ld_join(z, 'mc_out').astype(np.float64)
    shim = bm.rule_shim()
    out = np.zeros(3)
    import ctypes
    n_far = 0
    for row, w in zip(z['mc_in'], want):
        cl = shim.mdh_minimal_change(*[float(v) for v in row], out.ctypes.data_as(ctypes.c_void_p))
        assert bool(cl) == (not (w[1] == 1 and w[2] == 0))
        assert abs(out[2] - w[2]) <= 1e-12 * max(1.0, abs(w[2]))       # same minimum
        if abs(out[0] - w[0]) > 1e-12 * abs(w[0]) or abs(out[1] - w[1]) > 1e-9:
            n_far += 1                                                   # degenerate ties only
        out2 = np.zeros(3)                                               # two-thread split used by k_diag_block
        cl2 = shim.mdh_minimal_change_split(*[float(v) for v in row], out2.ctypes.data_as(ctypes.c_void_p))
        assert cl2 == cl and (np.array_equal(out, out2) or (np.isnan(out).any() and np.isnan(out2).any()))
    assert n_far <= 4


def test_upper_clamp_shortcut_equals_full_candidate_list():
    """md_upper_clamp_shortcut (csrc/md_rule.h) skips the two cubic solves when d* lies above [a, b] and b is
    max_diag_D.  Wherever it fires it must return what the reference's full candidate list returns: checked on the
    reference's golden _minimal_change vectors and on 20000 random upper-clamp situations (benign and huge alpha)."""
    import ctypes
    shim = bm.rule_shim()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    z = load_golden('scalar_rule.npz')
    want = ld_join(z, 'mc_out').astype(np.float64)
    quick, full = np.zeros(3), np.zeros(3)
    n_hit = 0
    for row, w in zip(z['mc_in'], want):
        args = [float(v) for v in row]
        if shim.mdh_upper_clamp_shortcut(*args, P(quick)):
            n_hit += 1
            assert shim.mdh_minimal_change_full(*args, P(full)) == 1
            assert quick[0] == full[0] and abs(quick[1] - full[1]) <= 1e-15 and abs(quick[2] - full[2]) <= 1e-15 * abs(full[2])
            assert abs(quick[0] - w[0]) <= 1e-15 * abs(w[0]) and abs(quick[1] - w[1]) <= 1e-12   # the reference itself
    rng = np.random.default_rng(0)
    n_rand = n_cheap = n_row = 0
    for _ in range(20000):
        scale = 10.0 ** rng.uniform(-3, 6)
        max_diag_D = scale * rng.uniform(0.5, 2.0)
        min_diag_D = max_diag_D * rng.choice([0.0, 1e-3, 0.2, 0.999, 1.0])
        alpha = scale * 10.0 ** rng.uniform(-6, 3)
        gamma = alpha + max_diag_D * (1 + 10.0 ** rng.uniform(-8, 2))        # d* above max_diag_D
        beta = scale ** 2 * 10.0 ** rng.uniform(-4, 6) * rng.choice([0.0, 1.0, 1.0])
        max_diag_B = rng.choice([np.inf, max_diag_D + alpha * rng.uniform(1.0, 3.0), max_diag_D + alpha])
        min_diag_B = rng.choice([-np.inf, 0.0, min_diag_D * 0.5])
        mav = rng.choice([3.3e-10, scale * 1e-3])
        args = [alpha, beta, gamma, min_diag_D, max_diag_D, min_diag_B, max_diag_B, mav]
        hit = shim.mdh_upper_clamp_shortcut(*args, P(quick))
        cheap = shim.mdh_upper_clamp_shortcut_cheap(*args)
        assert hit or not cheap, args        # the division-free test of the kernel's chain is a sufficient condition
        n_cheap += int(cheap)
        row = shim.mdh_upper_clamp_shortcut_row(*args)
        assert cheap or not row, args      # ... and so is its per-row precomputed form
        n_row += int(row)
        if not hit:
            continue
        n_rand += 1
        assert shim.mdh_minimal_change_full(*args, P(full)) == 1
        assert quick[0] == full[0], args
        # the full FP64 list may return a cubic root a hair below 1 whose f ties with f(b, 1) in double precision (the
        # u + v form of roots.py is only good to ~1e-12 when alpha << |d* - b|; the true root lies above 1 and is
        # clipped to omega = 1): the shortcut's omega = 1 is the exact answer, the difference is far below 1e-10
        assert abs(quick[1] - full[1]) <= 1e-9, (args, quick, full)
        assert abs(quick[2] - full[2]) <= 1e-12 * abs(full[2]), (args, quick, full)
    assert n_hit >= 20 and n_rand >= 10000, (n_hit, n_rand)
    assert n_cheap >= 0.95 * n_rand, (n_cheap, n_rand)     # ... that covers the benign regime
    assert n_row >= 0.95 * n_rand, (n_row, n_rand)
    # outside its range window the cheap test must stay silent (the full shortcut decides there)
    for alpha in (1e-120, 1e120, np.inf, np.nan):
        assert not shim.mdh_upper_clamp_shortcut_cheap(alpha, 1.0, alpha + 10.0, 0.1, 2.0, -np.inf, np.inf, 3.3e-10)
    assert not shim.mdh_upper_clamp_shortcut_cheap(1.0, np.inf, 11.0, 0.1, 2.0, -np.inf, np.inf, 3.3e-10)
    assert shim.mdh_upper_clamp_shortcut_cheap(1.0, 1.0, 11.0, 0.1, 2.0, -np.inf, np.inf, 3.3e-10)
    assert shim.mdh_upper_clamp_shortcut_row(1.0, 1.0, 11.0, 0.1, 2.0, -np.inf, np.inf, 3.3e-10)
    for alpha in (1e-120, 1e120, np.inf, np.nan):
        assert not shim.mdh_upper_clamp_shortcut_row(alpha, 1.0, alpha + 10.0, 0.1, 2.0, -np.inf, np.inf, 3.3e-10)
    # a gap between lo and max_diag_D at rounding level: the row form must not fire where the cheap form does not
    for g in (1e-9, 1e-8, 1e-7):
        args = (1.0, 1.0, 11.0, 2.0 * (1 - g), 2.0, -np.inf, np.inf, 3.3e-10)
        assert shim.mdh_upper_clamp_shortcut_cheap(*args) or not shim.mdh_upper_clamp_shortcut_row(*args)


def test_fp64_cubic_against_reference_vectors():
    z = load_golden('scalar_rule.npz')
    want = ld_join(z, 'cubic_out').astype(np.float64)
    import ctypes
    r = np.zeros(3)
    for (p, q), w in zip(z['cubic_in'], want):
        k = bm.rule_shim().mdh_cubic(p, q, r.ctypes.data_as(ctypes.c_void_p))
        w = w[~np.isnan(w)]
        assert k == len(w)
        np.testing.assert_allclose(np.sort(r[:k]), np.sort(w), rtol=2e-11, atol=1e-13)  # u + v cancels: last-bit differences in D are amplified


def test_cubic_overflow_rescaling():
    import ctypes
    r = np.zeros(3)
    p, q = 3e150, -2.9e150     # D overflows in FP64; root ~ -q/p
    k = bm.rule_shim().mdh_cubic(p, q, r.ctypes.data_as(ctypes.c_void_p))
    assert k == 1
    assert abs(r[0] ** 3 + p * r[0] + q) <= 1e-10 * abs(q)


CASES = [
    # (family, n, seed, s/lo-hi, kwargs, benign, nb)
    ('shifted_goe', 300, 1, 1.05, None, True, 64),
    ('shifted_goe', 300, 1, 1.05, None, True, 128),
    ('shifted_goe', 500, 2, 1.05, None, True, 128),
    ('shifted_goe', 300, 3, 0.9, None, False, 64),
    ('goe', 200, 4, None, dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values'), False, 64),
    ('goe', 200, 5, None, dict(min_diag_D=0.5, max_diag_D=3.0, permutation='increasing_absolute_diagonal_values'),
     False, 32),
    ('fixture', 96, 0, None, dict(min_diag_D=None, min_diag_B=1.0, max_diag_B=30.0, permutation='none'), False, 32),
    ('fixture', 96, 0, None, dict(min_diag_D=0.0, max_diag_D=50.0, permutation='decreasing_diagonal_values'), False, 32),
    ('lowrank', 200, 6, (-0.02, 1.0), dict(min_diag_D=1e-3, permutation='decreasing_diagonal_values'), True, 64),
]


def _make(family, n, seed, par, kw):
    if family == 'shifted_goe':
        A, kw2 = workloads.shifted_goe(n, seed, par)
        return A, (kw or kw2)
    if family == 'goe':
        return workloads.goe(n, seed), kw
    if family == 'fixture':
        return workloads.reference_fixture(n), kw
    if family == 'lowrank':
        return workloads.lowrank_plus_diag(n, seed, *par), kw
    raise ValueError(family)


@pytest.mark.parametrize('family,n,seed,par,kw,benign,nb', CASES)
def test_blocked_model_matches_oracle(family, n, seed, par, kw, benign, nb):
    A, kw = _make(family, n, seed, par, kw)
    ref = orc.approximate(A, **kw)
    kw2 = {k: v for k, v in kw.items() if k != 'permutation'}
    got = bm.blocked_factor(A, ref['p'], nb=nb, **kw2)
    np.testing.assert_array_equal(got['clamped'], ref['clamped'])
    d_ref = ref['d'].astype(np.float64)
    L_ref = ref['L']
    B_ref = L_ref @ (d_ref[:, None] * L_ref.T)
    B = got['L'] @ (got['d'][:, None] * got['L'].T)
    assert np.linalg.norm(B - B_ref) <= 1e-12 * np.linalg.norm(B_ref)
    if benign:
        np.testing.assert_allclose(got['d'], d_ref, rtol=1e-10)
        assert np.all(np.abs(got['L'] - L_ref) <= 1e-10 * np.maximum(1.0, np.abs(L_ref)))
        np.testing.assert_allclose(got['omega'], ref['omega'].astype(np.float64), rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(got['delta'], ref['delta'].astype(np.float64), rtol=1e-9,
                                   atol=1e-11 * np.abs(A).max())


def test_blocked_model_decompose_matches_potrf():
    A = workloads.lowrank_plus_diag(300, 5)
    ref = orc.decompose(A, 'decreasing_diagonal_values', 'LDL')
    got = bm.blocked_factor(A, ref['p'], rule=bm.RULE_EXACT, nb=64)
    assert got['info'] == 0 and not got['clamped'].any()
    np.testing.assert_allclose(got['d'], ref['d'], rtol=1e-11)
    np.testing.assert_allclose(got['L'], ref['L'], rtol=1e-10, atol=1e-13)
    # failing pivot index like LAPACK's info
    B = workloads.reference_fixture(40)
    _, info = orc.potrf_lower(B)
    got = bm.blocked_factor(B, np.arange(40), rule=bm.RULE_EXACT, nb=16)
    assert got['info'] == info > 0


@pytest.mark.parametrize('method', ['maximal_stability', 'minimal_difference'])
@pytest.mark.parametrize('family,n,seed,nb', [('spd', 160, 7, 32), ('indefinite', 120, 8, 48)])
def test_dynamic_pivoting_with_delayed_updates_matches_oracle(method, family, n, seed, nb):
    """The algorithm of csrc/kernels_dynamic.cuh (arg-min per column, symmetric swaps that carry the pending panel
    operands, pivot column = stored Schur value + pending updates, one product per panel) against the oracle's
    column-by-column restatement of Reimer.py:497-523."""
    if family == 'spd':
        A = workloads.goe(n, seed) + 1.05 * np.sqrt(2 * n) * np.eye(n)
        kw = dict(min_diag_D=0.2, max_diag_D=0.9 * 1.05 * np.sqrt(2 * n), max_diag_B=2.1 * np.sqrt(2 * n))
    else:
        A = workloads.goe(n, seed) + 2.0 * np.eye(n)
        kw = dict(min_diag_D=0.3, max_diag_D=30.0)
    ref = orc.approximate(A, permutation=method, **kw)
    got = bm.dynamic_delayed_factor(A, method, nb=nb, **kw)
    d_ref = ref['d'].astype(np.float64)
    if np.array_equal(got['p'], ref['p']):
        np.testing.assert_array_equal(got['clamped'], ref['clamped'])
        B_ref = ref['L'] @ (d_ref[:, None] * ref['L'].T)
        B = got['L'] @ (got['d'][:, None] * got['L'].T)
        assert np.linalg.norm(B - B_ref) <= 1e-12 * np.linalg.norm(B_ref)
        assert got['clamped'].any() and n // nb >= 2
    else:       # FP64 vs the reference's long double: the first divergence has to be a tie of the key
        i = int(np.argmax(got['p'] != ref['p']))
        assert abs(got['d'][i] - d_ref[i]) <= 1e-8 * abs(d_ref[i])
