"""Generate golden vectors from the UNMODIFIED Python reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    PYTHONOPTIMIZE=1 PYTHONDONTWRITEBYTECODE=1 python tests/golden/generate_goldens.py

* The two-line shim restores names that numpy/scipy removed: np.float and
  scipy.linalg.misc._datacopied.  The reference sources are not edited.
* PYTHONOPTIMIZE=1 is required: the assertion at Reimer.py:72 fires on ordinary inputs.
* `_minimal_change` is wrapped (not modified) to record the per-column f value, from which the
  "clamped" mask is derived (clamped <=> the call left the branch at Reimer.py:68 <=> omega != 1 or f > 0).

Outputs (committed): tests/golden/*.npz.  float128 vectors are stored as a float64 pair (hi, lo).
"""
import itertools
import json
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, '/root/reference')

# --- shim ---
np.float = float
import scipy.linalg.misc as _m  # noqa: E402
from scipy.linalg._misc import _datacopied  # noqa: E402
_m._datacopied = _datacopied

import matrix  # noqa: E402
import matrix.approximation.positive_semidefinite as aps  # noqa: E402
import matrix.approximation.positive_semidefinite.Reimer as R  # noqa: E402
import matrix.permute  # noqa: E402
import matrix._util.roots  # noqa: E402
import workloads  # noqa: E402

assert not __debug__, 'run with PYTHONOPTIMIZE=1'

_calls = []
_orig_minimal_change = R._minimal_change


def _recording_minimal_change(*args, **kwargs):
    out = _orig_minimal_change(*args, **kwargs)
    _calls.append(out)
    return out


R._minimal_change = _recording_minimal_change


def split(x):
    x = np.asarray(x, dtype=np.longdouble)
    hi = x.astype(np.float64)
    lo = (x - hi.astype(np.longdouble)).astype(np.float64)
    return hi, lo


def run_approx(A, kwargs, store, key, store_L=True, probes=None):
    """Run the reference; returns False when it raised."""
    _calls.clear()
    kw = dict(kwargs)
    n = A.shape[0]
    dynamic = kw.get('permutation', None) is None or (isinstance(kw.get('permutation'), str)
                                                      and kw['permutation'] in aps.APPROXIMATION_ONLY_PERMUTATION_METHODS)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dec = aps.decomposition(A.copy(), **kw)
    except Exception as e:  # ValueError for bad bounds etc.
        store[key + '/error'] = np.array(type(e).__name__)
        return False
    L = np.asarray(dec.L)
    d = np.asarray(dec.d)
    if dynamic:
        # the chosen (d, omega, f) per step is the one whose d/omega match the stored result
        f = np.full(n, np.nan, dtype=np.longdouble)
        cl = np.zeros(n, dtype=bool)
        idx = 0
        for i in range(n):
            block = _calls[idx: idx + (n - i)]
            idx += n - i
            # the same selection key as Reimer.py:499-508
            method = kw.get('permutation') or 'maximal_stability'
            if method == 'minimal_difference':
                best = min(enumerate(block), key=lambda t: (t[1][2], -t[1][0], t[1][1], t[0]))[1]
            else:
                best = min(enumerate(block), key=lambda t: (-t[1][0], t[1][2], t[1][1], t[0]))[1]
            assert best[0] == d[i] and best[1] == dec.omega[dec.p[i]]
            f[i] = best[2]
            cl[i] = not (best[1] == 1 and best[2] == 0)
    else:
        assert len(_calls) == n
        f = np.array([c[2] for c in _calls], dtype=np.longdouble)
        cl = np.array([not (c[1] == 1 and c[2] == 0) for c in _calls], dtype=bool)
    store[key + '/p'] = np.asarray(dec.p).astype(np.int64)
    store[key + '/d_hi'], store[key + '/d_lo'] = split(d)
    store[key + '/omega_hi'], store[key + '/omega_lo'] = split(dec.omega)
    store[key + '/delta_hi'], store[key + '/delta_lo'] = split(dec.delta)
    store[key + '/f_hi'], store[key + '/f_lo'] = split(f)
    store[key + '/clamped'] = cl
    if store_L:
        store[key + '/L'] = L
    if probes is not None:
        # B v = L D L^T v in permuted order, float64 products (checksum of the composed matrix)
        d64 = d.astype(np.float64)
        store[key + '/probes'] = probes
        store[key + '/Bv'] = L @ (d64[:, None] * (L.T @ probes))
        store[key + '/Lv'] = L @ probes
        rows = np.unique(np.linspace(0, n - 1, 8).astype(int))
        store[key + '/L_rows_idx'] = rows
        store[key + '/L_rows'] = L[rows]
    return True


def kw_to_json(kw):
    out = {}
    for k, v in kw.items():
        if isinstance(v, np.ndarray):
            out[k] = v.tolist()
        else:
            out[k] = v
    return out


def gen_approx_fixture10():
    """The parametrisation grid of matrix/tests/test_approximate.py on its own n=10 fixture."""
    n = 10
    A = workloads.reference_fixture(n)
    store = {'A': A}
    manifest = []
    perms = list(matrix.UNIVERSAL_PERMUTATION_METHODS) + list(aps.APPROXIMATION_ONLY_PERMUTATION_METHODS) + \
        [workloads.reference_permutation(n)]
    grid = itertools.product((None, 1, np.arange(n) / n), (None, 1, np.arange(n) + 1), (None, 0, 1), (None, 1),
                             (None, 0.01), perms)
    for idx, (mB, MB, mD, MD, mavD, perm) in enumerate(grid):
        kw = dict(min_diag_B=mB, max_diag_B=MB, min_diag_D=mD, max_diag_D=MD, min_abs_value_D=mavD, permutation=perm)
        key = f'c{idx:04d}'
        ok = run_approx(A, kw, store, key)
        manifest.append(dict(key=key, kwargs=kw_to_json(kw), ok=ok))
    store['manifest'] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(HERE, 'approx_fixture10.npz'), **store)
    print('approx_fixture10:', len(manifest), 'cases,', sum(m['ok'] for m in manifest), 'ok')


def gen_approx_families():
    store = {}
    manifest = []
    rng = np.random.default_rng(7)

    def add(name, family, A, kw, store_L, note=''):
        n = A.shape[0]
        probes = rng.standard_normal((n, 3))
        ok = run_approx(A, kw, store, name, store_L=store_L, probes=probes)
        if A.shape[0] <= 64:
            store[name + '/A'] = A
        manifest.append(dict(key=name, family=family, n=n, kwargs=kw_to_json(kw), ok=ok, note=note))
        print(name, 'ok' if ok else 'RAISED', 'clamped', int(store[name + '/clamped'].sum()) if ok else '-')

    A = np.array([[2.0, 1.0], [1.0, -3.0]])
    add('known2x2', dict(kind='literal'), A, dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values'), True,
        'known answer quoted in SURVEY 8c')
    add('goe64', dict(kind='goe', n=64, seed=64), workloads.goe(64, 64),
        dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values'), True, 'explosive regime')
    for s in (1.05, 0.9):
        A, kw = workloads.shifted_goe(200, 200, s)
        add(f'f2_n200_s{s}', dict(kind='shifted_goe', n=200, seed=200, s=s), A, kw, True,
            'benign' if s > 1 else 'explosive: compare clamped set and B only')
    A = workloads.lowrank_plus_diag(128, 128, -0.02, 1.0)
    add('f3_n128_nearpsd', dict(kind='lowrank_plus_diag', n=128, seed=128, lo=-0.02, hi=1.0), A,
        dict(min_diag_D=1e-3, permutation='decreasing_absolute_diagonal_values'), True)
    add('f3_n128_nearpsd_default', dict(kind='lowrank_plus_diag', n=128, seed=128, lo=-0.02, hi=1.0), A,
        dict(min_diag_D=None, min_diag_B=0.05, permutation='increasing_diagonal_values'), True)
    A = workloads.reference_fixture(96)
    add('fix96_maxstab', dict(kind='reference_fixture', n=96), A, dict(min_diag_D=0.5, max_diag_D=400.0,
                                                                       permutation='maximal_stability'), True)
    add('fix96_mindiff', dict(kind='reference_fixture', n=96), A, dict(min_diag_D=0.5, max_diag_B=60.0,
                                                                       permutation='minimal_difference'), True)
    # above the size where L is stored: decisions + probes only
    A, kw = workloads.shifted_goe(1000, 1000, 1.05)
    add('f2_n1000_s1.05', dict(kind='shifted_goe', n=1000, seed=1000, s=1.05), A, kw, False, 'benign')
    A, kw = workloads.shifted_goe(600, 600, 0.9)
    add('f2_n600_s0.9', dict(kind='shifted_goe', n=600, seed=600, s=0.9), A, kw, False, 'explosive')
    A = workloads.goe(2000, 2000)
    add('c1_goe2000', dict(kind='goe', n=2000, seed=2000), A,
        dict(min_diag_D=1e-6, permutation='decreasing_diagonal_values'), False,
        'BASELINE config 1; explosive: compare clamped set and B only')
    store['manifest'] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(HERE, 'approx_families.npz'), **store)


def gen_scalar_rule():
    """Known-answer vectors for _minimal_change and solver_depressed_cubic."""
    rng = np.random.default_rng(11)
    rows = []
    outs = []
    for _ in range(4000):
        alpha = float(10 ** rng.uniform(-3, 3)) if rng.random() > 0.1 else 0.0
        beta = float(alpha * 10 ** rng.uniform(-2, 2)) if alpha > 0 else 0.0
        gamma = float(rng.normal(0, 5))
        mD = float(rng.choice([0.0, 1e-6, 0.1, 1.0]))
        MD = float(rng.choice([math.inf, 1.0, 5.0]))
        mB = float(rng.choice([-math.inf, 0.0, 0.5, 1.0]))
        MB = float(rng.choice([math.inf, 1.0, 8.0]))
        mav = float(rng.choice([3.3e-10, 0.01]))
        if max(mD, mav, mB) > min(MD, MB):
            continue
        try:
            d, w, f = _orig_minimal_change(np.longdouble(alpha), np.longdouble(beta), gamma, mD, max_diag_D=MD,
                                           min_diag_B=mB, max_diag_B=MB, min_abs_value_D=mav)
        except Exception:
            continue
        rows.append([alpha, beta, gamma, mD, MD, mB, MB, mav])
        outs.append([np.longdouble(d), np.longdouble(w), np.longdouble(f)])
    outs = np.array(outs, dtype=np.longdouble)
    hi, lo = split(outs)
    cub_in = []
    cub_out = []
    for _ in range(2000):
        p = float(rng.normal(0, 3))
        q = float(rng.normal(0, 3))
        r = matrix._util.roots.solver_depressed_cubic(np.longdouble(p), np.longdouble(q), include_complex_values=False)
        r = [np.longdouble(x) for x in r]
        cub_in.append([p, q])
        cub_out.append(r + [np.longdouble(np.nan)] * (3 - len(r)))
    chi, clo = split(np.array(cub_out, dtype=np.longdouble))
    np.savez_compressed(os.path.join(HERE, 'scalar_rule.npz'), mc_in=np.array(rows), mc_out_hi=hi, mc_out_lo=lo,
                        cubic_in=np.array(cub_in), cubic_out_hi=chi, cubic_out_lo=clo)
    print('scalar_rule:', len(rows), 'minimal_change vectors,', len(cub_in), 'cubic vectors')


def gen_decompose_solve_permute():
    store = {}
    manifest = []
    n = 10
    A = workloads.reference_fixture_spd(n)
    store['spd10/A'] = A
    perms = list(matrix.UNIVERSAL_PERMUTATION_METHODS) + [workloads.reference_permutation(n)]
    for pi, perm in enumerate(perms):
        for rt in matrix.DECOMPOSITION_TYPES:
            key = f'spd10/p{pi}_{rt}'
            dec = matrix.decompose(A.copy(), permutation=perm, return_type=rt)
            store[key + '/p'] = np.asarray(dec.p).astype(np.int64)
            if rt == 'LL':
                store[key + '/L'] = np.asarray(dec.L)
            elif rt == 'LDL':
                store[key + '/L'] = np.asarray(dec.L)
                store[key + '/d'] = np.asarray(dec.d, dtype=np.float64)
            else:
                store[key + '/LD'] = np.asarray(dec.LD)
            b1 = workloads.reference_fixture(n)[:, 0]
            b2 = workloads.reference_fixture(n)[:, :2]
            store[key + '/x1'] = np.asarray(dec.solve(b1), dtype=np.float64)
            store[key + '/x2'] = np.asarray(dec.solve(b2), dtype=np.float64)
            store[key + '/Ab2'] = np.asarray(dec.matrix_right_side_multiplication(b2), dtype=np.float64)
            manifest.append(dict(key=key, n=n, permutation=perm.tolist() if isinstance(perm, np.ndarray) else perm,
                                 return_type=rt))
    store['spd10/b1'] = b1
    store['spd10/b2'] = b2
    # larger SPD (F3) LDL with permutation + solve
    A = workloads.lowrank_plus_diag(128, 5)
    dec = matrix.decompose(A.copy(), permutation='decreasing_diagonal_values', return_type='LDL')
    rng = np.random.default_rng(3)
    b = rng.standard_normal((128, 5))
    store['f3_128/L'] = np.asarray(dec.L)
    store['f3_128/d'] = np.asarray(dec.d, dtype=np.float64)
    store['f3_128/p'] = np.asarray(dec.p).astype(np.int64)
    store['f3_128/b'] = b
    store['f3_128/x'] = np.asarray(dec.solve(b))
    store['f3_128/x0'] = np.asarray(dec.solve(b[:, 0]))
    # failure: indefinite matrix -> NoDecompositionPossibleWithProblematicSubdecompositionError
    A = workloads.reference_fixture(10)
    try:
        matrix.decompose(A.copy(), permutation='none')
        raise SystemExit('expected failure')
    except matrix.errors.NoDecompositionPossibleWithProblematicSubdecompositionError as e:
        store['indef10/A'] = A
        store['indef10/bad_index'] = np.array(e.problematic_leading_principal_submatrix_index)
    # solve of an approximate() result (float128 d)
    A, kw = workloads.shifted_goe(200, 200, 1.05)
    dec = aps.decomposition(A.copy(), **kw)
    b = rng.standard_normal((200, 3))
    store['f2_200_solve/b'] = b
    store['f2_200_solve/x'] = np.asarray(dec.solve(b), dtype=np.float64)
    # permute
    A = workloads.reference_fixture(10)
    p = workloads.reference_permutation(10)
    store['permute10/A'] = A
    store['permute10/p'] = p.astype(np.int64)
    store['permute10/B'] = matrix.permute.symmetric(A, p)
    for m in matrix.UNIVERSAL_PERMUTATION_METHODS:
        store[f'permvec10/{m}'] = np.asarray(matrix.permute.permutation_vector(A, m)).astype(np.int64)
    Ag = workloads.goe(300, 1)
    for m in matrix.UNIVERSAL_PERMUTATION_METHODS:
        store[f'permvec300/{m}'] = np.asarray(matrix.permute.permutation_vector(Ag, m)).astype(np.int64)
    # positive_semidefinite_matrix (B) on the fixture
    A = workloads.reference_fixture(10)
    for ci, kw in enumerate([dict(min_diag_D=1.0, permutation='none'),
                             dict(min_diag_D=0, max_diag_B=(np.arange(10) + 1.0), permutation='decreasing_diagonal_values'),
                             dict(min_diag_B=1.0, max_diag_D=1.0, permutation='maximal_stability')]):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            B = aps.positive_semidefinite_matrix(A.copy(), **kw)
        store[f'psd10/c{ci}/B'] = np.asarray(B, dtype=np.float64)
        manifest.append(dict(key=f'psd10/c{ci}', kwargs=kw_to_json(kw)))
    store['manifest'] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(HERE, 'decompose_solve_permute.npz'), **store)
    print('decompose_solve_permute done')


if __name__ == '__main__':
    which = sys.argv[1:] or ['scalar', 'fixture10', 'families', 'dsp']
    if 'scalar' in which:
        gen_scalar_rule()
    if 'dsp' in which:
        gen_decompose_solve_permute()
    if 'fixture10' in which:
        gen_approx_fixture10()
    if 'families' in which:
        gen_approx_families()
