"""The CPU oracle (oracle/mdo_oracle.c) against golden vectors produced by the unmodified Python
reference (tests/golden/generate_goldens.py).  This is what pins the oracle."""
import json

import numpy as np
import pytest

import workloads
from oracle import cpu_oracle as orc
from conftest import load_golden, ld_join


def _kw_from_json(kw):
    out = {}
    for k, v in kw.items():
        out[k] = np.asarray(v, dtype=float) if isinstance(v, list) and k != 'permutation' else v
        if k == 'permutation' and isinstance(v, list):
            out[k] = np.asarray(v, dtype=np.int64)
    return out


def test_cubic_against_reference():
    z = load_golden('scalar_rule.npz')
    want = ld_join(z, 'cubic_out')
    for (p, q), w in zip(z['cubic_in'], want):
        got = orc.cubic_real_roots(p, q)
        w = w[~np.isnan(w)]
        assert len(got) == len(w)
        np.testing.assert_allclose(np.array(got, dtype=np.longdouble), w, rtol=2e-15, atol=2e-15)  # libm cbrt vs scipy (cephes) cbrt differ by a few ulp


def test_minimal_change_against_reference():
    z = load_golden('scalar_rule.npz')
    want = ld_join(z, 'mc_out')
    bad = 0
    for row, w in zip(z['mc_in'], want):
        alpha, beta, gamma, mD, MD, mB, MB, mav = row
        d, om, f, cl = orc.minimal_change(alpha, beta, gamma, mD, MD, mB, MB, mav)
        assert d == w[0] or abs(d - w[0]) <= 1e-15 * abs(w[0])
        assert abs(om - w[1]) <= 4e-15
        assert abs(f - w[2]) <= 1e-14 * max(1.0, abs(w[2]))
        assert cl == (not (w[1] == 1 and w[2] == 0))
    assert bad == 0


def _check_case(z, key, A, kw, tight=True, stats=None):
    """p and the clamped mask must be identical.  d / omega / delta / L must agree unless the reference's
    own choice was a rounding-noise tie (two candidates with f equal to ~1e-16): then f must
    agree at the tie and the composed matrix B = L D L^T must still match."""
    res = orc.approximate(A, **kw)
    np.testing.assert_array_equal(res['p'], z[key + '/p'])
    np.testing.assert_array_equal(res['clamped'], z[key + '/clamped'])
    d = ld_join(z, key + '/d')
    om = ld_join(z, key + '/omega')
    de = ld_join(z, key + '/delta')
    f = ld_join(z, key + '/f')
    rt = 1e-13 if tight else 1e-9
    bad = np.nonzero(np.abs(res['d'] - d) > rt * np.abs(d))[0]
    if len(bad):
        i = bad[0]
        assert abs(res['f'][i] - f[i]) <= 1e-14 * abs(f[i]), 'not a tie'
        if stats is not None:
            stats['ties'] = stats.get('ties', 0) + 1
        if key + '/L' in z:
            Lr = z[key + '/L']
            Br = Lr @ (d.astype(float)[:, None] * Lr.T)
            B = res['L'] @ (res['d'].astype(float)[:, None] * res['L'].T)
            assert np.linalg.norm(B - Br) <= 1e-12 * np.linalg.norm(Br)
        return res
    np.testing.assert_allclose(res['omega'], om, rtol=rt, atol=1e-300)
    np.testing.assert_allclose(res['delta'], de, rtol=1e-9, atol=1e-12 * max(1.0, float(np.abs(A).max())))
    np.testing.assert_allclose(res['f'], f, rtol=1e-12, atol=1e-300)
    if key + '/L' in z:
        np.testing.assert_allclose(res['L'], z[key + '/L'], rtol=rt, atol=1e-14 if tight else 1e-10)
    return res


def test_approximate_reference_test_grid_n10():
    """Every parametrisation of matrix/tests/test_approximate.py (dense, real) that the reference
    completes on its own fixture, static and dynamic permutations."""
    z = load_golden('approx_fixture10.npz')
    A = z['A']
    np.testing.assert_array_equal(A, workloads.reference_fixture(10))
    manifest = json.loads(str(z['manifest']))
    n_ok = 0
    stats = {}
    for m in manifest:
        if not m['ok']:
            continue
        _check_case(z, m['key'], A, _kw_from_json(m['kwargs']), stats=stats)
        n_ok += 1
    assert n_ok > 800
    assert stats.get('ties', 0) <= 16  # documented near-tie class (iii)


@pytest.mark.parametrize('key', ['known2x2', 'goe64', 'f2_n200_s1.05', 'f2_n200_s0.9', 'f3_n128_nearpsd',
                                 'f3_n128_nearpsd_default', 'fix96_maxstab', 'fix96_mindiff'])
def test_approximate_families(key):
    z = load_golden('approx_families.npz')
    m = {e['key']: e for e in json.loads(str(z['manifest']))}[key]
    A = _family_matrix(z, m)
    _check_case(z, key, A, _kw_from_json(m['kwargs']), tight=key not in ('goe64', 'f2_n200_s0.9'))


def _family_matrix(z, m):
    fam = m['family']
    if fam['kind'] == 'literal':
        return z[m['key'] + '/A']
    if fam['kind'] == 'goe':
        return workloads.goe(fam['n'], fam['seed'])
    if fam['kind'] == 'shifted_goe':
        return workloads.shifted_goe(fam['n'], fam['seed'], fam['s'])[0]
    if fam['kind'] == 'lowrank_plus_diag':
        return workloads.lowrank_plus_diag(fam['n'], fam['seed'], fam['lo'], fam['hi'])
    if fam['kind'] == 'reference_fixture':
        return workloads.reference_fixture(fam['n'])
    raise ValueError(fam)


@pytest.mark.parametrize('key', ['f2_n1000_s1.05', 'f2_n600_s0.9', 'c1_goe2000'])
def test_approximate_large_decisions_and_probes(key):
    z = load_golden('approx_families.npz')
    m = {e['key']: e for e in json.loads(str(z['manifest']))}[key]
    A = _family_matrix(z, m)
    tight = 's1.05' in key
    res = _check_case(z, key, A, _kw_from_json(m['kwargs']), tight=tight)  # c1_goe2000 = BASELINE config 1
    L = res['L']
    d64 = res['d'].astype(np.float64)
    V = z[key + '/probes']
    Bv = L @ (d64[:, None] * (L.T @ V))
    np.testing.assert_allclose(Bv, z[key + '/Bv'], rtol=1e-10, atol=1e-10 * np.abs(z[key + '/Bv']).max())
    np.testing.assert_allclose(L[z[key + '/L_rows_idx']], z[key + '/L_rows'], rtol=1e-12 if tight else 1e-6,
                               atol=1e-13 if tight else 1e-8)


def test_known_answer_2x2():
    """SURVEY 8c known answer."""
    res = orc.approximate(np.array([[2.0, 1.0], [1.0, -3.0]]), min_diag_D=1e-6,
                          permutation='decreasing_diagonal_values')
    np.testing.assert_array_equal(res['p'], [0, 1])
    np.testing.assert_allclose(res['d'].astype(float), [2, 1e-6], rtol=1e-15)
    np.testing.assert_allclose(res['omega'].astype(float), [1, 0.3938888], rtol=1e-6)
    np.testing.assert_allclose(res['delta'].astype(float), [0, 3.07757519], rtol=1e-7, atol=1e-15)
    np.testing.assert_allclose(res['L'][1, 0], 0.1969444, rtol=1e-6)


def test_decompose_solve_permute():
    z = load_golden('decompose_solve_permute.npz')
    manifest = json.loads(str(z['manifest']))
    A = z['spd10/A']
    np.testing.assert_allclose(A, workloads.reference_fixture_spd(10), rtol=0, atol=0)
    for m in manifest:
        key = m['key']
        if not key.startswith('spd10/'):
            continue
        rt = m['return_type']
        perm = m['permutation']
        perm = np.asarray(perm, dtype=np.int64) if isinstance(perm, list) else perm
        res = orc.decompose(A, perm, 'LL' if rt == 'LL' else 'LDL')
        assert res['info'] == 0
        np.testing.assert_array_equal(res['p'], z[key + '/p'])
        if rt == 'LL':
            np.testing.assert_allclose(res['L'], z[key + '/L'], rtol=1e-12, atol=1e-14)
        elif rt == 'LDL':
            np.testing.assert_allclose(res['L'], z[key + '/L'], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(res['d'], z[key + '/d'], rtol=1e-12)
        else:
            LD = res['L'].copy()
            np.fill_diagonal(LD, res['d'])
            np.testing.assert_allclose(LD, z[key + '/LD'], rtol=1e-12, atol=1e-14)
        x1 = orc.solve(res['L'], res['d'], res['p'], z['spd10/b1'])
        x2 = orc.solve(res['L'], res['d'], res['p'], z['spd10/b2'])
        np.testing.assert_allclose(x1, z[key + '/x1'], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(x2, z[key + '/x2'], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(orc.multiply(res['L'], res['d'], res['p'], z['spd10/b2']), z[key + '/Ab2'],
                                   rtol=1e-12, atol=1e-12)
    # F3 n=128
    A = workloads.lowrank_plus_diag(128, 5)
    res = orc.decompose(A, 'decreasing_diagonal_values', 'LDL')
    np.testing.assert_array_equal(res['p'], z['f3_128/p'])
    np.testing.assert_allclose(res['L'], z['f3_128/L'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(res['d'], z['f3_128/d'], rtol=1e-12)
    np.testing.assert_allclose(orc.solve(res['L'], res['d'], res['p'], z['f3_128/b']), z['f3_128/x'], rtol=1e-10,
                               atol=1e-12)
    np.testing.assert_allclose(orc.solve(res['L'], res['d'], res['p'], z['f3_128/b'][:, 0]), z['f3_128/x0'],
                               rtol=1e-10, atol=1e-12)
    # failure index
    res = orc.decompose(z['indef10/A'], 'none', 'LL')
    assert res['info'] - 1 == int(z['indef10/bad_index'])
    # permutation
    np.testing.assert_array_equal(orc.permute_symmetric(z['permute10/A'], z['permute10/p']), z['permute10/B'])
    for meth in ('none', 'decreasing_diagonal_values', 'increasing_diagonal_values',
                 'decreasing_absolute_diagonal_values', 'increasing_absolute_diagonal_values'):
        np.testing.assert_array_equal(orc.permutation_vector(z['permute10/A'], meth), z[f'permvec10/{meth}'])
        np.testing.assert_array_equal(orc.permutation_vector(workloads.goe(300, 1), meth), z[f'permvec300/{meth}'])


def test_psd_matrix():
    z = load_golden('decompose_solve_permute.npz')
    manifest = json.loads(str(z['manifest']))
    A = workloads.reference_fixture(10)
    for m in manifest:
        if not m['key'].startswith('psd10/'):
            continue
        kw = _kw_from_json(m['kwargs'])
        res = orc.approximate(A, **kw)
        B = orc.psd_matrix(A, res, kw.get('min_diag_B'), kw.get('max_diag_B'))
        np.testing.assert_allclose(B, z[m['key'] + '/B'], rtol=1e-13, atol=1e-13)


def test_solve_approximate_result():
    z = load_golden('decompose_solve_permute.npz')
    A, kw = workloads.shifted_goe(200, 200, 1.05)
    res = orc.approximate(A, **kw)
    x = orc.solve(res['L'], res['d'].astype(np.float64), res['p'], z['f2_200_solve/b'])
    np.testing.assert_allclose(x, z['f2_200_solve/x'], rtol=1e-10, atol=1e-12 * np.abs(x).max())


def test_fp64_twin_matches_in_benign_regime():
    A, kw = workloads.shifted_goe(300, 9, 1.05)
    ref = orc.approximate(A, **kw)
    kw2 = {k: v for k, v in kw.items() if k != 'permutation'}
    twin = orc.approximate_f64(A, ref['p'], **kw2)
    np.testing.assert_array_equal(twin['clamped'], ref['clamped'])
    np.testing.assert_allclose(twin['L'], ref['L'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(twin['d'], ref['d'].astype(float), rtol=1e-12)


def test_gmw_oracle_against_reference_goldens():
    """oracle/gmw_oracle.py (numpy restatement of GMW_SE.py) on the 96 cases produced by the unmodified reference
    (tests/golden/generate_gmw_goldens.py): GMW_81 / GMW_T1 / GMW_T2 x bounds x matrix families, bit for bit."""
    import json
    from oracle import gmw_oracle
    z = load_golden('gmw.npz')
    manifest = json.loads(str(z['manifest']))
    assert len(manifest) == 96 and all(m['ok'] for m in manifest)
    for m in manifest:
        A = z[m['matrix'] + '/A']
        kw = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m['kwargs'].items()}
        B = gmw_oracle.approximate(A, m['fn'], **kw)
        np.testing.assert_array_equal(B, z[m['key'] + '/B'], err_msg=m['key'])


def test_se_oracle_against_reference_goldens():
    """The SE half of oracle/gmw_oracle.py on the 120 cases of tests/golden/se.npz (SE_90 / SE_99 / SE_T1 x bounds x
    matrix families, n = 1 ... 96), produced by the reference with the size-1-array -> scalar conversion of older numpy
    restored at GMW_SE.py:180 (tests/golden/generate_se_goldens.py): bit for bit.  SE_T2 is SE_99 (GMW_SE.py:558)."""
    import json
    from oracle import gmw_oracle
    z = load_golden('se.npz')
    manifest = json.loads(str(z['manifest']))
    assert len(manifest) == 120 and all(m['ok'] for m in manifest)
    for m in manifest:
        A = z[m['matrix'] + '/A']
        kw = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m['kwargs'].items()}
        B = gmw_oracle.approximate(A, m['fn'], **kw)
        np.testing.assert_array_equal(B, z[m['key'] + '/B'], err_msg=m['key'])
        if m['fn'] == 'SE_99':
            np.testing.assert_array_equal(gmw_oracle.approximate(A, 'SE_T2', **kw), B)
