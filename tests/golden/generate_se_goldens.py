"""Golden vectors of the SE family (matrix/approximation/positive_semidefinite/GMW_SE.py: SE_90, SE_99, SE_T1; SE_T2 is
an alias of SE_99, :558) from the Python reference.  Build container only:

    PYTHONOPTIMIZE=1 PYTHONDONTWRITEBYTECODE=1 python tests/golden/generate_se_goldens.py

With numpy >= 1.25 every SE_* call of the unmodified reference fails at the last-but-one column: GMW_SE.py:180 builds
np.array([[a, c], [c, A]], dtype=float64) from a scalar a, a length-1 vector c and a 1 x 1 matrix A, which older numpy
versions accepted by turning the size-1 arrays into scalars and current ones reject ("inhomogeneous shape").  The
reference's sources are not touched: the module is imported as it is and ONLY its name `np` is replaced by a proxy
whose `array` retries a failed construction with the size-1 arrays of the nested list converted to scalars -- the
conversion numpy used to do by itself.  No arithmetic of the reference is changed.
Output (committed): tests/golden/se.npz."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, '/root/reference')
np.float = float
import scipy.linalg.misc as _m  # noqa: E402
from scipy.linalg._misc import _datacopied  # noqa: E402
_m._datacopied = _datacopied
import matrix.approximation.positive_semidefinite.GMW_SE as G  # noqa: E402
import workloads  # noqa: E402

assert not __debug__, 'run with PYTHONOPTIMIZE=1'


def _scalars(obj):
    if isinstance(obj, (list, tuple)):
        return [_scalars(o) for o in obj]
    if isinstance(obj, np.ndarray) and obj.size == 1:
        return obj.reshape(()).item()
    return obj


class _NumpyWithOldScalarConversion:
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(obj, *args, **kwargs):
        try:
            return np.array(obj, *args, **kwargs)
        except ValueError:
            return np.array(_scalars(obj), *args, **kwargs)


G.np = _NumpyWithOldScalarConversion()


def main():
    out, manifest = {}, []
    cases = []
    rng = np.random.default_rng(17)
    for n in (1, 2, 3, 12, 40, 96):
        g = rng.standard_normal((n, n))
        cases.append(('goe%d' % n, (g + g.T) / 2))
    cases.append(('f2_90_s0.9', workloads.shifted_goe(90, 3, 0.9)[0]))
    cases.append(('spd60', workloads.lowrank_plus_diag(60, 4)))
    cases.append(('spd2', np.array([[2.0, 0.5], [0.5, 1.0]])))
    cases.append(('fixture10', workloads.reference_fixture(10)))
    for name, A in cases:
        out[name + '/A'] = A
        n = A.shape[0]
        kwlist = [dict(), dict(min_diag_D=1e-3), dict(min_diag_B=0.5 * np.abs(A).max(), max_diag_B=3.0 * np.abs(A).max()),
                  dict(max_diag_B=np.linspace(1.0, 2.0, n) * np.abs(A).max())]
        for fn in ('SE_90', 'SE_99', 'SE_T1'):
            for ik, kw in enumerate(kwlist):
                key = '%s/%s/%d' % (name, fn, ik)
                entry = dict(key=key, matrix=name, fn=fn,
                             kwargs={k: (v.tolist() if isinstance(v, np.ndarray) else float(v)) for k, v in kw.items()})
                try:
                    B = getattr(G, fn)(A.copy(), **kw)
                    out[key + '/B'] = np.asarray(B, dtype=np.float64)
                    entry['ok'] = True
                except Exception as e:  # noqa: BLE001
                    entry['ok'] = False
                    entry['error'] = type(e).__name__
                manifest.append(entry)
    out['manifest'] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(HERE, 'se.npz'), **out)
    print(len(manifest), 'cases,', sum(m['ok'] for m in manifest), 'ok')
    print([m['key'] + ':' + m.get('error', '') for m in manifest if not m['ok']][:10])


if __name__ == '__main__':
    main()
