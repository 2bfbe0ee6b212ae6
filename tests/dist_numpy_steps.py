"""numpy stand-in for the per-rank device steps of matrix_decomposition_b200.distributed (test
infrastructure): same buffer layout and call protocol as CudaSteps / mdb_dist_*, arithmetic taken from
oracle/blocked_model.py.  Lets the panel loop (ownership, slot rotation, lookahead ordering, broadcast
protocol) run on CPU over gloo."""
import math

import numpy as np
import torch

from oracle import blocked_model as bm
from matrix_decomposition_b200.distributed import NB, local_cols, local_blocks, panel_elems, owner_of


class GlooComm:
    def broadcast(self, tensor, src, stream):
        import torch.distributed as dist
        work = dist.broadcast(tensor, src=src, async_op=True)

        class Handle:
            def wait(self, s):
                work.wait()
        return Handle()


class NumpySteps:
    def __init__(self, n, rank, world, Ap, rule=bm.RULE_REIMER, min_diag_D=None, max_diag_D=None, min_abs_value_D=None,
                 min_abs_value_L=None, min_diag_B=-math.inf, max_diag_B=math.inf, agg=2):
        self.n, self.rank, self.world = n, rank, world
        self.ld = local_cols(n, rank, world)
        self.nloc = local_blocks(n, rank, world)
        self.A = np.zeros((n, self.ld))
        for lb in range(self.nloc):
            g0 = (lb * world + rank) * NB
            w = min(NB, n - g0)
            self.A[:, lb * NB: lb * NB + w] = Ap[:, g0:g0 + w]
        self.S = self.A.copy()
        self.gamma = Ap.diagonal().copy()
        self.alpha, self.beta, self.rowmax = np.zeros(n), np.zeros(n), np.zeros(n)
        self.d, self.omega, self.delta = np.zeros(n), np.zeros(n), np.zeros(n)
        self.clamped = np.zeros(n, dtype=bool)
        self.rule = rule
        self.mD = math.nan if min_diag_D is None else float(min_diag_D)
        self.MD = math.inf if max_diag_D is None else float(max_diag_D)
        self.mavD = float(np.finfo(np.longdouble).eps ** 0.5) if min_abs_value_D is None else float(min_abs_value_D)
        self.mavL = float(np.finfo(np.float64).eps) if min_abs_value_L is None else float(min_abs_value_L)
        self.mB, self.MB = float(min_diag_B), float(max_diag_B)
        self.panels = [torch.zeros(panel_elems(n, 0), dtype=torch.float64) for _ in range(2)]
        self.agg = agg
        self.OPX = [np.zeros((n, agg * NB)) for _ in range(2)]      # operands of one outer block, rows = global - ob
        self.OPY = [np.zeros((n, agg * NB)) for _ in range(2)]
        self.log = []

    def outer_width(self, k):
        T = (self.n + NB - 1) // NB
        return max(1, min(self.agg, T - k))

    def panel_tensor(self, slot, k):
        return self.panels[slot][:panel_elems(self.n, k)]

    def _views(self, slot, k):
        n = self.n
        k1 = min((k + 1) * NB, n)
        rows = n - k1
        buf = self.panels[slot].numpy()
        hdr = buf[:5 * NB].reshape(5, NB)
        vec = buf[5 * NB:5 * NB + 3 * rows].reshape(3, rows)
        MX = buf[5 * NB + 3 * rows: 5 * NB + 3 * rows + rows * NB].reshape(rows, NB)
        return hdr, vec, MX

    def factor_panel(self, k, slot, k_outer, oslot, stream):
        assert owner_of(k, self.world) == self.rank
        self.log.append(('factor', k, slot, k_outer, oslot))
        n = self.n
        k0, k1 = k * NB, min((k + 1) * NB, n)
        kb = k1 - k0
        lc0 = (k // self.world) * NB
        hdr, vec, MX = self._views(slot, k)
        X = np.zeros((n - k0, kb))
        Lt = np.zeros((kb, kb))
        for c in range(kb):
            K = k0 + c
            dk, wk, _, cl = bm.decide(self.rule, self.mD, self.MD, self.mavD, self.alpha[K], self.beta[K], self.gamma[K],
                                      self.mB, self.MB)
            self.d[K], self.omega[K], self.clamped[K] = dk, wk, cl
            self.delta[K] = (dk - self.gamma[K]) + (wk * wk * self.alpha[K] if wk != 0 else 0.0)
            drop = (wk == 0) or (wk * self.rowmax[K] < self.mavL)
            hdr[0, c], hdr[1, c], hdr[2, c], hdr[3, c], hdr[4, c] = dk, wk, float(drop), self.delta[K], float(cl)
            if wk != 0:
                t = wk * X[c, :c]
                t[np.abs(t) < self.mavL] = 0
                Lt[c, :c] = t
            u = self.d[k0:K] * Lt[c, :c]
            a = self.A[K + 1:, lc0 + c]
            r = a.copy() if drop else (1 - wk) * a + wk * self.S[K + 1:, lc0 + c]
            acc = np.zeros(n - K - 1)
            for m in range(c):
                acc += X[c + 1:, m] * u[m]
            x = (r - acc) / dk if dk != 0 else np.zeros_like(r)
            X[c + 1:, c] = x
            self.alpha[K + 1:] += x * x * dk
            self.beta[K + 1:] += 2 * a * a
            self.rowmax[K + 1:] = np.maximum(self.rowmax[K + 1:], np.abs(x))
        for c in range(kb):
            self.S[k0 + c + 1:, lc0 + c] = X[c + 1:, c]
        if k1 < n:
            ob, c0 = k_outer * NB, k0 - k_outer * NB
            self.OPX[oslot][k1 - ob: n - ob, c0:c0 + NB] = X[kb:, :]
            self.OPY[oslot][k1 - ob: n - ob, c0:c0 + NB] = -X[kb:, :] * self.d[k0:k1][None, :]
            MX[:] = X[kb:, :]
            vec[0], vec[1], vec[2] = self.alpha[k1:], self.beta[k1:], self.rowmax[k1:]

    def absorb_panel(self, k, slot, k_outer, oslot, stream):
        self.log.append(('absorb', k, slot, k_outer, oslot))
        if owner_of(k, self.world) == self.rank:
            return
        n = self.n
        k0, k1 = k * NB, min((k + 1) * NB, n)
        hdr, vec, MX = self._views(slot, k)
        kb = k1 - k0
        self.d[k0:k1], self.omega[k0:k1], self.delta[k0:k1] = hdr[0, :kb], hdr[1, :kb], hdr[3, :kb]
        self.clamped[k0:k1] = hdr[4, :kb] != 0
        if k1 < n:
            self.alpha[k1:], self.beta[k1:], self.rowmax[k1:] = vec[0], vec[1], vec[2]
            ob, c0 = k_outer * NB, k0 - k_outer * NB
            self.OPX[oslot][k1 - ob: n - ob, c0:c0 + NB] = MX
            self.OPY[oslot][k1 - ob: n - ob, c0:c0 + NB] = -MX * hdr[0][None, :]

    def apply_update(self, k_outer, oslot, k_first, k_count, blk_lo, blk_hi, stream):
        self.log.append(('apply', k_outer, oslot, k_first, k_count, blk_lo, blk_hi))
        n = self.n
        T = (n + NB - 1) // NB
        if blk_hi < 0 or blk_hi > T:
            blk_hi = T
        blk_lo = max(blk_lo, k_first + k_count)
        r0, ob = (k_first + k_count) * NB, k_outer * NB
        if r0 >= n:
            return
        c0, c1 = (k_first - k_outer) * NB, (k_first + k_count - k_outer) * NB
        PX = self.OPX[oslot][r0 - ob: n - ob, c0:c1]
        PY = self.OPY[oslot][r0 - ob: n - ob, c0:c1]
        for lb in range(self.nloc):
            J = lb * self.world + self.rank
            if J < blk_lo or J >= blk_hi:
                continue
            g0 = J * NB
            w = min(NB, n - g0)
            upd = PX @ PY[g0 - r0: g0 - r0 + w].T          # rows r0.., columns g0..g0+w
            rows = np.arange(r0, n)[:, None]
            cols = np.arange(g0, g0 + w)[None, :]
            blk = self.S[r0:, lb * NB: lb * NB + w]
            blk += np.where(rows > cols, upd, 0.0)

    def local_L(self):
        n = self.n
        L = np.zeros((n, self.ld))
        for lb in range(self.nloc):
            g0 = (lb * self.world + self.rank) * NB
            for c in range(min(NB, n - g0)):
                gc = g0 + c
                col = self.omega[gc + 1:] * self.S[gc + 1:, lb * NB + c]
                col[np.abs(col) < self.mavL] = 0
                L[gc + 1:, lb * NB + c] = col
                L[gc, lb * NB + c] = 1.0
        return L
