"""CPU restatement (numpy, test infrastructure only) of the GMW family of modified Cholesky approximations of the
reference: matrix/approximation/positive_semidefinite/GMW_SE.py.  Pinned against tests/golden/gmw.npz, which was
produced by the unmodified reference (tests/golden/generate_gmw_goldens.py).

  _modified_LDLt   GMW_SE.py:6-35    right-looking rank-1 LDL^T without pivoting (j = 0, :17-18), d_k = choose_d(...)
  _modified_A      GMW_SE.py:38-64   B = A + diag(delta), optional diagonal rescaling to [min_diag_B, max_diag_B]
  _approximate     GMW_SE.py:67-118  two-phase strategy around choose_d
  _GMW             GMW_SE.py:121-152 beta^2 and d = |c|_inf^2 / beta^2
  GMW_81 / GMW_T1 / GMW_T2  GMW_SE.py:197-346  parameter sets
  _SE              GMW_SE.py:155-194 d = |c|_1, the 2 x 2 eigenvalue step for the last two columns
  SE_90 / SE_99 / SE_T1 (SE_T2 = SE_99)  GMW_SE.py:374-558  parameter sets
The SE family is pinned against tests/golden/se.npz (tests/golden/generate_se_goldens.py: the reference run with a numpy
proxy that restores the size-1-array -> scalar conversion numpy >= 1.25 dropped, GMW_SE.py:180)."""
import numpy as np

EPS = float(np.finfo(np.float64).eps)

VARIANTS = {   # GMW_SE.py:243, :293, :343
    'GMW_81': dict(use_two_phases=False, nondecreasing=False, use_abs=True, mu=None),
    'GMW_T1': dict(use_two_phases=True, nondecreasing=True, use_abs=True, mu=0.75),
    'GMW_T2': dict(use_two_phases=True, nondecreasing=True, use_abs=False, mu=0.75),
}


def gmw_delta(A, use_two_phases, nondecreasing, use_abs, mu, min_diag_D=None):
    """d and delta of _modified_LDLt with the choose_d of _approximate / _GMW."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    if min_diag_D is None:
        min_diag_D = EPS ** 0.5                                           # :77-80
    phase = 1 if use_two_phases else 1.5                                  # :84
    if mu is not None:                                                    # :87-93
        min_d_next = -mu * np.abs(A.diagonal()).max() if n else 0.0
        min_d_factor = -mu
    else:
        min_d_next, min_d_factor = min_diag_D, -np.inf
    previous_delta = 0.0
    beta_squared = None
    S = A.copy()
    d = np.empty(n)
    delta = np.empty(n)
    for k in range(n):
        a = S[k, k]
        c = S[k + 1:, k].copy()
        T = S[k + 1:, k + 1:]                                             # the matrix choose_d sees (:24-26)
        if phase == 1:                                                    # :100-104
            if a >= min_diag_D and np.all(T.diagonal() - c ** 2 / a >= min_d_next) and np.all(T.diagonal() >= min_d_factor * a):
                dk = a
            else:
                phase = 1.5
        if phase == 1.5:                                                  # :106-108 -> _GMW.init_choose_d_state :123-142
            m = T.shape[0]
            if m > 1:
                max_diag = np.abs(T.diagonal()).max()
                max_off = np.abs(np.tril(T, -1)).max()
                x = m ** 2 - 1 if use_abs else m ** 2 - m
                beta_squared = max(EPS, max_diag, max_off / x ** 0.5)
            elif m == 1:
                beta_squared = max(EPS, abs(T[0, 0]))
            else:
                beta_squared = EPS
            phase = 2
        if phase == 2:                                                    # :110-117, _GMW.choose_d :145-148
            dk = (np.abs(c).max() ** 2 / beta_squared) if len(c) > 0 else 0.0
            dk = max(dk, min_diag_D)
            a_changed = abs(a) if use_abs else a
            if nondecreasing:
                a_changed += previous_delta
            dk = max(dk, a_changed)
            previous_delta = dk - a
        d[k] = dk
        delta[k] = dk - a                                                 # :29 (p is the identity)
        S[k + 1:, k + 1:] -= np.outer(c, c) / dk                          # :31
    return d, delta


SE_VARIANTS = {   # GMW_SE.py:430, :493, :555, :558
    'SE_90': dict(use_two_phases=True, nondecreasing=True, use_abs=False, mu=None),
    'SE_99': dict(use_two_phases=True, nondecreasing=True, use_abs=False, mu=0.1),
    'SE_T1': dict(use_two_phases=True, nondecreasing=False, use_abs=True, mu=0.1),
    'SE_T2': dict(use_two_phases=True, nondecreasing=True, use_abs=False, mu=0.1),
}


def se_delta(A, use_two_phases, nondecreasing, use_abs, mu, min_diag_D=None):
    """d and delta of _modified_LDLt with the choose_d of _approximate / _SE (GMW_SE.py:155-192)."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    tau = EPS ** (2 / 3) if mu is not None else EPS ** (1 / 3)            # :158-161
    if min_diag_D is None:
        min_diag_D = EPS ** 0.5
    phase = 1 if use_two_phases else 1.5
    if mu is not None:
        min_d_next = -mu * np.abs(A.diagonal()).max() if n else 0.0
        min_d_factor = -mu
    else:
        min_d_next, min_d_factor = min_diag_D, -np.inf
    previous_delta = 0.0
    tau_max = None
    d_last = None
    S = A.copy()
    d = np.empty(n)
    delta = np.empty(n)
    for k in range(n):
        a = S[k, k]
        c = S[k + 1:, k].copy()
        T = S[k + 1:, k + 1:]
        if phase == 1:
            if a >= min_diag_D and np.all(T.diagonal() - c ** 2 / a >= min_d_next) and np.all(T.diagonal() >= min_d_factor * a):
                dk = a
            else:
                phase = 1.5
        if phase == 1.5:                                                  # init_choose_d_state, :163-169
            tau_max = tau * (np.abs(T.diagonal()).max() if len(T) > 0 else 0)
            phase = 2
        if phase == 2:                                                    # choose_d, :171-190
            m = len(c)
            if m > 1:
                dk = max(np.linalg.norm(c, ord=1), tau_max)
            elif m == 1:
                l1, l2 = np.sort(np.linalg.eigvalsh(np.array([[a, c[0]], [c[0], T[0, 0]]], dtype=np.float64)))
                dk = a - l1 + max(tau * (l2 - l1) / (1 - tau), tau_max)
                if use_abs:
                    dk = max(dk, a - 2 * l1)
                d_last = dk
            else:
                if d_last is not None:
                    dk = d_last
                else:
                    dk = max((-EPS ** (1 / 3) * a) / (1 - EPS ** (1 / 3)), tau_max)
            dk = max(dk, min_diag_D)                                      # :111-117
            a_changed = abs(a) if use_abs else a
            if nondecreasing:
                a_changed += previous_delta
            dk = max(dk, a_changed)
            previous_delta = dk - a
        d[k] = dk
        delta[k] = dk - a
        S[k + 1:, k + 1:] -= np.outer(c, c) / dk
    return d, delta


def modified_matrix(A, delta, min_diag_B=None, max_diag_B=None):
    """_modified_A, GMW_SE.py:43-62."""
    B = np.array(A, dtype=np.float64)
    n = B.shape[0]
    B[np.diag_indices(n)] += delta
    if min_diag_B is not None or max_diag_B is not None:
        lo = -np.inf if min_diag_B is None else min_diag_B
        hi = np.inf if max_diag_B is None else max_diag_B
        old = B.diagonal().copy()
        new = np.minimum(np.maximum(old, lo), hi)
        s = (new / old) ** 0.5
        B = (s[:, None] * B) * s[None, :]
        B[np.diag_indices(n)] = new
    return B


def approximate(A, fn, min_diag_B=None, max_diag_B=None, min_diag_D=None):
    if fn in SE_VARIANTS:
        d, delta = se_delta(A, min_diag_D=min_diag_D, **SE_VARIANTS[fn])
    else:
        d, delta = gmw_delta(A, min_diag_D=min_diag_D, **VARIANTS[fn])
    return modified_matrix(A, delta, min_diag_B, max_diag_B)
