"""CPU-side tests: the C-ABI library loads and exports every symbol of include/mdb200.h, the host
mirror validates arguments like the reference (same exception types, before any device work), the
bookkeeping of the decomposition objects matches the reference's semantics, and there is no CPU
fallback: compute calls fail loudly without a CUDA device."""
import json
import os
import re
import tempfile

import numpy as np
import pytest

import workloads
import matrix_decomposition_b200 as matrix
from matrix_decomposition_b200 import _native
from conftest import load_golden, REPO


def _have_gpu():
    try:
        _native.context()
        return True
    except matrix.errors.NativeLibraryError:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, 'include', 'mdb200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(mdb_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 40
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert lib.mdb_version() >= 100


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'matrix_decomposition_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            src = open(os.path.join(root, f), errors='ignore').read() if f.endswith(('.py', '.cu', '.cuh', '.h')) else ''
            if f.endswith('.py'):
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'libmdo_oracle' not in src and 'host_rule.so' not in src, f
            elif src:
                assert not re.search(r'#include\s+"[^"]*oracle', src) and 'dlopen' not in src, f


@pytest.mark.skipif(_have_gpu(), reason='a CUDA device is present')
def test_no_cpu_fallback_without_a_gpu():
    A = workloads.reference_fixture_spd(10)
    with pytest.raises(matrix.errors.NativeLibraryError):
        matrix.decompose(A)
    with pytest.raises(matrix.errors.NativeLibraryError):
        matrix.approximation.positive_semidefinite.decomposition(A, permutation='none')
    with pytest.raises(matrix.errors.NativeLibraryError):
        matrix.permute.symmetric(A, np.arange(10)[::-1])
    dec = matrix.decompositions.LL_Decomposition(np.linalg.cholesky(A))
    with pytest.raises(matrix.errors.NativeLibraryError):
        dec.solve(np.ones(10))


def test_argument_validation_matches_reference_error_cases():
    """Every parametrisation of the reference's own test grid that made the reference raise ValueError
    raises ValueError here too, on the host, before any device work."""
    z = load_golden('approx_fixture10.npz')
    A = z['A']
    manifest = json.loads(str(z['manifest']))
    n_err = 0
    for m in manifest:
        if m['ok']:
            continue
        kw = {k: (np.asarray(v, dtype=float) if isinstance(v, list) and k != 'permutation' else v)
              for k, v in m['kwargs'].items()}
        if isinstance(kw['permutation'], list):
            kw['permutation'] = np.asarray(kw['permutation'])
        assert str(z[m['key'] + '/error']) == 'ValueError'
        with pytest.raises(ValueError):
            matrix.approximation.positive_semidefinite.decomposition(A, **kw)
        n_err += 1
    assert n_err == 36
    aps = matrix.approximation.positive_semidefinite
    with pytest.raises(matrix.errors.MatrixNotSquareError):
        aps.decomposition(np.ones((3, 4)))
    with pytest.raises(ValueError):
        aps.decomposition(A, return_type='XX', permutation='none')
    with pytest.raises(ValueError):
        aps.decomposition(A, permutation='DUMMY_METHOD')
    with pytest.raises(ValueError):
        aps.decomposition(A, permutation=np.arange(5))
    with pytest.raises(ValueError):
        aps.decomposition(A, min_diag_D=-1.0, permutation='none')
    with pytest.raises(ValueError):
        aps.decomposition(A, min_diag_B=np.ones(3), permutation='none')
    with pytest.raises(ValueError):
        aps.decomposition(A, min_diag_D=2.0, max_diag_D=1.0, permutation='none')
    with pytest.raises(matrix.errors.MatrixNotSquareError):
        matrix.decompose(np.ones((3, 4)))
    with pytest.raises(matrix.errors.MatrixNotFiniteError):
        matrix.decompose(np.array([[1.0, np.nan], [np.nan, 1.0]]))
    with pytest.raises(ValueError):
        matrix.decompose(np.eye(3), permutation='DUMMY_METHOD')
    with pytest.raises(ValueError):
        matrix.decompose(np.eye(3), return_type='XX')


def test_permutation_vectors_match_reference():
    z = load_golden('decompose_solve_permute.npz')
    A10 = z['permute10/A']
    A300 = workloads.goe(300, 1)
    for meth in matrix.UNIVERSAL_PERMUTATION_METHODS:
        np.testing.assert_array_equal(matrix.permute.permutation_vector(A10, meth), z[f'permvec10/{meth}'])
        np.testing.assert_array_equal(matrix.permute.permutation_vector(A300, meth), z[f'permvec300/{meth}'])
    with pytest.raises(ValueError):
        matrix.permute.permutation_vector(A10, 'DUMMY_METHOD')
    p = workloads.reference_permutation(10)
    q = matrix.permute.invert_permutation_vector(p)
    np.testing.assert_array_equal(p[q], np.arange(10))
    np.testing.assert_array_equal(q[p], np.arange(10))
    assert matrix.permute.symmetric(A10, None) is A10


def test_decomposition_objects_bookkeeping():
    """Conversions, equality, copy semantics, save/load -- host-side, same semantics as the reference
    (tests/test_convert.py, test_equal.py, test_io.py, test_append_block_decomposition.py)."""
    z = load_golden('decompose_solve_permute.npz')
    key = 'spd10/p1_'
    p = z[key + 'LDL/p']
    ldl = matrix.decompositions.LDL_Decomposition(z[key + 'LDL/L'], z[key + 'LDL/d'], p)
    ll = ldl.as_type('LL')
    comp = ldl.as_type('LDL_compressed')
    np.testing.assert_allclose(ll.L, z[key + 'LL/L'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(comp.LD, z[key + 'LDL_compressed/LD'], rtol=1e-12, atol=1e-14)
    back = ll.as_type('LDL')
    assert back.is_almost_equal(ldl) and not back.is_equal(ll)
    assert comp.as_type('LDL').is_equal(ldl)
    assert ldl.as_type('LDL') is ldl and ldl.as_type(None) is ldl and ldl.as_type('LDL', copy=True) is not ldl
    assert ldl.copy().is_equal(ldl) and ldl == ldl.copy()
    # setters copy (decompositions.py:842) and check the unit diagonal (:843-844)
    Lsrc = z[key + 'LDL/L'].copy()
    d2 = matrix.decompositions.LDL_Decomposition(Lsrc, z[key + 'LDL/d'], p)
    Lsrc[3, 1] = 99
    assert d2.L[3, 1] != 99
    with pytest.raises(ValueError):
        matrix.decompositions.LDL_Decomposition(2 * np.eye(3), np.ones(3))
    # features
    assert ldl.is_finite() and ldl.is_positive_definite() and ldl.is_invertible() and not ldl.is_singular()
    sing = matrix.decompositions.LDL_Decomposition(np.eye(3), np.array([1.0, 0.0, 2.0]))
    assert sing.is_singular() and not sing.is_positive_definite() and sing.is_positive_semidefinite()
    with pytest.raises(matrix.errors.DecompositionSingularError):
        sing.check_invertible()
    with pytest.raises(matrix.errors.DecompositionSingularError):
        sing.solve(np.ones(3))   # raised on the host, before any device work (decompositions.py:985)
    neg = matrix.decompositions.LDL_Decomposition(np.eye(3), np.array([1.0, -1.0, 2.0]), np.array([2, 0, 1]))
    with pytest.raises(matrix.errors.NoDecompositionPossibleWithProblematicSubdecompositionError) as ei:
        neg.as_type('LL')
    assert ei.value.problematic_leading_principal_submatrix_index == 0
    # permutation properties
    assert ldl.is_permuted and not matrix.decompositions.LL_Decomposition(np.eye(4)).is_permuted
    np.testing.assert_array_equal(ldl.p_inverse[ldl.p], np.arange(10))
    assert ldl.P.shape == (10, 10)
    # append
    both = ldl.append_block_decomposition(ll)
    assert both.n == 20 and both.type_str == 'LDL'
    np.testing.assert_array_equal(both.p[10:], p + 10)
    # save / load (tar of npz, constants.py:31-35)
    with tempfile.TemporaryDirectory() as tmp:
        for dec in (ldl, ll, comp):
            fn = os.path.join(tmp, 'x_' + dec.type_str)
            dec.save(fn)
            assert os.path.exists(fn + '.dec')
            again = matrix.decompositions.load(fn)
            assert again.is_equal(dec)
        with pytest.raises(matrix.errors.DecompositionInvalidDecompositionTypeFile):
            matrix.decompositions.LL_Decomposition().load(os.path.join(tmp, 'x_LDL'))


def test_constants_and_facade_names():
    assert matrix.DECOMPOSITION_TYPES == ('LDL', 'LDL_compressed', 'LL')
    assert matrix.UNIVERSAL_PERMUTATION_METHODS[0] == 'none'
    aps = matrix.approximation.positive_semidefinite
    assert aps.APPROXIMATION_ONLY_PERMUTATION_METHODS == ('minimal_difference', 'maximal_stability')
    with pytest.warns(DeprecationWarning):
        fn = matrix.decomposition
    assert fn is aps.decomposition
    for name in ('decompose', 'is_positive_definite', 'is_positive_semidefinite', 'is_invertible', 'solve',
                 'decompositions', 'errors', 'approximation', 'permute'):
        assert hasattr(matrix, name)
    matrix.release_device_memory()   # no context yet: nothing to do, and no attempt to create one


def test_gmw_family_host_side_contract():
    """GMW_SE.py surface: argument errors are raised on the host before any device work; SE_T2 is SE_99
    (GMW_SE.py:558)."""
    aps = matrix.approximation.positive_semidefinite
    for name in ('GMW_81', 'GMW_T1', 'GMW_T2', 'SE_90', 'SE_99', 'SE_T1', 'SE_T2'):
        assert callable(getattr(aps, name))
    with pytest.raises(matrix.errors.MatrixNotSquareError):
        aps.GMW_T1(np.ones((3, 4)))
    with pytest.raises(ValueError):
        aps.GMW_81(np.eye(3), min_diag_D=0.0)
    with pytest.raises(ValueError):
        aps.GMW_T2(np.eye(3) * 1j)
    assert aps.SE_T2 is aps.SE_99
    with pytest.raises(matrix.errors.MatrixNotSquareError):
        aps.SE_90(np.ones((2, 3)))
    with pytest.raises(ValueError):
        aps.SE_T1(np.eye(3), min_diag_D=-1.0)
