"""Golden vectors of the GMW family (matrix/approximation/positive_semidefinite/GMW_SE.py: GMW_81, GMW_T1, GMW_T2)
from the UNMODIFIED Python reference.  Build container only:

    PYTHONOPTIMIZE=1 PYTHONDONTWRITEBYTECODE=1 python tests/golden/generate_gmw_goldens.py

The SE family of the same file (SE_90, SE_99, SE_T1, SE_T2) cannot be pinned: with numpy 2.x every call fails at the
last-but-one column (GMW_SE.py:172, np.array([[a, c], [c, A]]) with array-valued c and A raises ValueError).
Output (committed): tests/golden/gmw.npz."""
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


def main():
    out, manifest = {}, []
    cases = []
    rng = np.random.default_rng(7)
    for n in (2, 3, 12, 40, 96):
        g = rng.standard_normal((n, n))
        cases.append(('goe%d' % n, (g + g.T) / 2))
    cases.append(('f2_90_s0.9', workloads.shifted_goe(90, 3, 0.9)[0]))
    cases.append(('spd60', workloads.lowrank_plus_diag(60, 4)))
    cases.append(('fixture10', workloads.reference_fixture(10)))
    for name, A in cases:
        out[name + '/A'] = A
        n = A.shape[0]
        kwlist = [dict(), dict(min_diag_D=1e-3), dict(min_diag_B=0.5 * np.abs(A).max(), max_diag_B=3.0 * np.abs(A).max()),
                  dict(max_diag_B=np.linspace(1.0, 2.0, n) * np.abs(A).max())]
        for fn in ('GMW_81', 'GMW_T1', 'GMW_T2'):
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
    np.savez_compressed(os.path.join(HERE, 'gmw.npz'), **out)
    print(len(manifest), 'cases,', sum(m['ok'] for m in manifest), 'ok')


if __name__ == '__main__':
    main()
