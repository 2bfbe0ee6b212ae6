"""Multi-GPU factorization: 1-D block-cyclic column distribution, one process per GPU.

Global column block J (width 128, in elimination order) lives on rank J % world.  Per panel k the owner
factors the diagonal block and the panel on its GPU (mdb_dist_factor_panel); the result -- the unscaled
rows X, the (d, omega, ...) header and the alpha / beta / rowmax slices -- is broadcast with
torch.distributed (NCCL over NVLink / NVSwitch) and absorbed by every rank into its operand buffers
(X and Y = -X D, mdb_dist_absorb_panel).  Updates use the same two-level blocking as the single-GPU path:
`outer_width` consecutive panels form an outer block; after panel k only the remaining column blocks of
its outer block are updated (K = 128), and everything beyond the outer block gets one bulk update per
outer block with K = 128 * width on the FP64 DMMA tile kernel (mdb_dist_apply_update), each rank on its own
column blocks.  Lookahead: the owner of the next panel updates that column block first, then factors and
broadcasts it while all ranks are still busy with the bulk update on a second stream, so the broadcast and
the sequential panel work hide under the update.

The orchestration below is backend-neutral: `Steps` is the per-rank device work, `Comm` the broadcast.
The product backend is CudaSteps (libmdb200.so + torch.cuda streams) with NcclComm; the CPU tests in
tests/test_distributed_cpu.py drive the same loop over gloo with a numpy stand-in for the device steps.
"""
import ctypes
import math
import os

import numpy as np

NB = 128


def owner_of(k, world):
    return k % world


def local_blocks(n, rank, world):
    nblk = (n + NB - 1) // NB
    return (nblk - rank + world - 1) // world if nblk > rank else 0


def local_cols(n, rank, world):
    return max(local_blocks(n, rank, world), 1) * NB


def panel_elems(n, k):
    k1 = min((k + 1) * NB, n)
    rows = n - k1
    return 5 * NB + 3 * rows + rows * NB


class NullStream:
    """Stream / event stand-in for CPU backends (everything is synchronous)."""

    def record(self):
        return None

    def wait(self, event):
        pass


def factorize(n, rank, world, steps, comm, main=None, bulk=None, lookahead=True):
    """The panel loop.  `steps` provides outer_width / factor_panel / absorb_panel / apply_update / panel_tensor,
    `comm.broadcast(tensor, src, stream)` returns a handle whose wait(stream) makes `stream` wait for it.
    `main` / `bulk` are stream objects with record() -> event and wait(event)."""
    main = main or NullStream()
    bulk = bulk or NullStream()
    T = (n + NB - 1) // NB
    if T == 0:
        return
    works = {}
    bulk_done = {}

    def issue_factor_and_broadcast(k, k_outer, oslot, outer_index):
        slot = k % 2
        if owner_of(k, world) == rank:
            steps.factor_panel(k, slot, k_outer, oslot, main)
        works[k] = comm.broadcast(steps.panel_tensor(slot, k), owner_of(k, world), main)

    issue_factor_and_broadcast(0, 0, 0, 0)
    kO, oidx = 0, 0
    while kO < T:
        cnt = steps.outer_width(kO)
        kE = kO + cnt
        oslot = oidx % 2
        for k in range(kO, kE):
            works.pop(k).wait(main)                                   # panel k has arrived
            steps.absorb_panel(k, k % 2, kO, oslot, main)
            if k + 1 < kE:                                            # inner updates: stay inside the outer block
                steps.apply_update(kO, oslot, k, 1, k + 1, k + 2, main)   # next column block first
                issue_factor_and_broadcast(k + 1, kO, oslot, oidx)
                steps.apply_update(kO, oslot, k, 1, k + 2, kE, main)
        if kE < T:
            cntn = steps.outer_width(kE)
            if oidx - 1 in bulk_done:
                # bulk update O-1 reached the next outer block and released the other operand buffer
                main.wait(bulk_done.pop(oidx - 1))
            if lookahead:
                ready = main.record()                                 # operands of this outer block complete
                steps.apply_update(kO, oslot, kO, cnt, kE, kE + 1, main)          # next panel's block first
                issue_factor_and_broadcast(kE, kE, (oidx + 1) % 2, oidx + 1)     # overlaps with the bulk update
                steps.apply_update(kO, oslot, kO, cnt, kE + 1, kE + cntn, main)   # rest of the next outer block
                bulk.wait(ready)
                steps.apply_update(kO, oslot, kO, cnt, kE + cntn, -1, bulk)       # everything beyond
                bulk_done[oidx] = bulk.record()
            else:
                steps.apply_update(kO, oslot, kO, cnt, kE, -1, main)
                bulk_done[oidx] = main.record()
                issue_factor_and_broadcast(kE, kE, (oidx + 1) % 2, oidx + 1)
        kO, oidx = kE, oidx + 1
    for ev in bulk_done.values():
        main.wait(ev)


# ------------------------------------------------------------------------------------------------
# product backend: libmdb200.so + torch.cuda + NCCL
# ------------------------------------------------------------------------------------------------
class TorchStream:
    def __init__(self, stream):
        self.stream = stream

    def record(self):
        import torch
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def wait(self, event):
        if event is not None:
            self.stream.wait_event(event)


def init_process_group(device_index=None, **kw):
    """torch.distributed.init_process_group('nccl') with NCCL's internal stream at HIGH priority.  The panel
    broadcast is on the critical path while the bulk update keeps every SM busy on a low-priority stream; at the
    default priority the broadcast kernel queues behind the ~10^4 CTAs of that update and the lookahead is lost."""
    import torch
    import torch.distributed as dist
    if device_index is None:
        device_index = torch.cuda.current_device()
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return dist.init_process_group('nccl', device_id=torch.device('cuda', device_index), pg_options=opts, **kw)


class NcclComm:
    def broadcast(self, tensor, src, stream):
        import torch
        import torch.distributed as dist
        with torch.cuda.stream(stream.stream):
            work = dist.broadcast(tensor, src=src, async_op=True)

        class Handle:
            def wait(self, s):
                with torch.cuda.stream(s.stream):
                    work.wait()
        return Handle()


class CudaSteps:
    """Per-rank device state: S_loc / A_loc / panel buffers are torch tensors, the steps are C-ABI calls."""

    def __init__(self, n, rank, world, device=None):
        import torch
        from . import _native
        self._native = _native
        self.n, self.rank, self.world = n, rank, world
        self.device = torch.cuda.current_device() if device is None else device
        self.ctx = _native.context(self.device)
        self.ctx.trim()     # the slabs below are torch tensors: give cached single-GPU buffers back first
        self.ld = local_cols(n, rank, world)
        dev = torch.device('cuda', self.device)
        self.S = torch.empty((n, self.ld), dtype=torch.float64, device=dev)
        self.A = torch.empty((n, self.ld), dtype=torch.float64, device=dev)
        pe = panel_elems(n, 0)
        self.panels = [torch.empty(pe, dtype=torch.float64, device=dev) for _ in range(2)]
        h = ctypes.c_void_p()
        rc = self.ctx.lib.mdb_dist_create(self.ctx.handle, n, rank, world, ctypes.c_void_p(self.S.data_ptr()),
                                          ctypes.c_void_p(self.A.data_ptr()), self.ld,
                                          ctypes.c_void_p(self.panels[0].data_ptr()),
                                          ctypes.c_void_p(self.panels[1].data_ptr()), ctypes.byref(h))
        self.ctx.check(rc, 'mdb_dist_create')
        self.handle = h
        self.bulk_events = None

    def close(self):
        if self.handle:
            self.ctx.lib.mdb_dist_destroy(self.handle)
            self.handle = None

    def info(self):
        """The device-side status word of the current factorization (mdb_dist_info)."""
        v = ctypes.c_int(0)
        self.ctx.check(self.ctx.lib.mdb_dist_info(self.handle, ctypes.byref(v)), 'mdb_dist_info')
        return int(v.value)

    def _on(self, stream):
        self.ctx.set_stream(stream.stream.cuda_stream)

    def generate_f2(self, seed, mu, p, stream):
        self._on(stream)
        p64 = None if p is None else np.ascontiguousarray(p, dtype=np.int64)
        self.ctx.check(self.ctx.lib.mdb_dist_generate_f2(self.handle, seed, mu, self._native.ptr(p64)), 'mdb_dist_generate_f2')

    def set_matrix(self, Ap, stream):
        self._on(stream)
        Ap = np.ascontiguousarray(Ap, dtype=np.float64)
        self.ctx.check(self.ctx.lib.mdb_dist_set_matrix(self.handle, self._native.ptr(Ap), Ap.shape[1]), 'mdb_dist_set_matrix')

    def set_bounds(self, bounds, p, stream):
        self._on(stream)
        p64 = None if p is None else np.ascontiguousarray(p, dtype=np.int64)
        self.ctx.check(self.ctx.lib.mdb_dist_set_bounds(self.handle, ctypes.byref(bounds), self._native.ptr(p64)),
                       'mdb_dist_set_bounds')

    def panel_tensor(self, slot, k):
        return self.panels[slot][:panel_elems(self.n, k)]

    def outer_width(self, k):
        return int(self.ctx.lib.mdb_dist_outer_width(self.handle, k))

    def factor_panel(self, k, slot, k_outer, oslot, stream):
        self._on(stream)
        self.ctx.check(self.ctx.lib.mdb_dist_factor_panel(self.handle, k, slot, k_outer, oslot), 'mdb_dist_factor_panel')

    def absorb_panel(self, k, slot, k_outer, oslot, stream):
        self._on(stream)
        self.ctx.check(self.ctx.lib.mdb_dist_absorb_panel(self.handle, k, slot, k_outer, oslot), 'mdb_dist_absorb_panel')

    def apply_update(self, k_outer, oslot, k_first, k_count, blk_lo, blk_hi, stream):
        self._on(stream)
        timed = self.bulk_events is not None and k_first == k_outer and k_count == self.outer_width(k_outer)
        if timed:       # serialised profiling pass: CUDA events around every bulk update on its launching stream
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream.stream)
        self.ctx.check(self.ctx.lib.mdb_dist_apply_update(self.handle, k_outer, oslot, k_first, k_count, blk_lo, blk_hi),
                       'mdb_dist_apply_update')
        if timed:
            e1.record(stream.stream)
            self.bulk_events.append((e0, e1))

    def set_local(self, A_loc_host, stream, lower_only=False):
        """This rank's column slab from a host array in slab layout (n x ld_loc).  lower_only: transfer per column block
        only the rows from its diagonal block downwards -- all the factorization, the solve and the gather read, half of
        the bytes; the identity / residual helpers below, which multiply by the whole original slab, then do not apply."""
        self._on(stream)
        assert A_loc_host.shape == (self.n, self.ld) and A_loc_host.dtype == np.float64 and A_loc_host.flags.c_contiguous
        fn = self.ctx.lib.mdb_dist_set_local_lower if lower_only else self.ctx.lib.mdb_dist_set_local
        self.ctx.check(fn(self.handle, self._native.ptr(A_loc_host), self.ld, self._native.HOST), 'mdb_dist_set_local')

    def finalize(self, stream, collective=True):
        """Finalises the local columns and returns the replicated vectors.  Raises on EVERY rank when any rank's
        device-side status word is non-zero (a failing pivot is only seen by the rank that owns its panel): the word is
        all-reduced (min and max) before the local status is looked at."""
        self._on(stream)
        n = self.n
        d, omega, delta = np.empty(n), np.empty(n), np.empty(n)
        clamped = np.zeros(n, dtype=np.uint8)
        P = self._native.ptr
        rc = self.ctx.lib.mdb_dist_finalize(self.handle, P(d), P(omega), P(delta), P(clamped))
        if collective and self.world > 1:
            check_status_all_ranks(self.info(), self.device)
        self.ctx.check(rc, 'mdb_dist_finalize')
        return d, omega, delta, clamped.astype(bool)


def check_status_all_ranks(info, device=None):
    """All-reduces a rank-local status word; raises the same NativeLibraryError on every rank when any is non-zero."""
    import torch
    import torch.distributed as dist
    from . import errors
    dev = torch.device('cuda', device) if (device is not None and dist.get_backend() == 'nccl') else torch.device('cpu')
    t = torch.tensor([info, -info], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hi, lo = int(t[0].item()), -int(t[1].item())
    if hi == 0 and lo == 0:
        return
    if lo < 0:
        raise errors.NativeLibraryError('distributed factorization: device-side status {} on some rank; the factor is not '
                                        'valid'.format(lo))
    raise errors.NativeLibraryError('distributed factorization: pivot {} is not positive'.format(hi - 1))


def _streams(device):
    import torch
    lo, hi = -1, 0
    main = TorchStream(torch.cuda.Stream(device=device, priority=-1))
    bulk = TorchStream(torch.cuda.Stream(device=device, priority=0))
    return main, bulk


def approximate_distributed(Ap_or_none, n, bounds_kwargs, p=None, generator=None, lookahead=True):
    """Factorizes over all ranks of the default process group (NCCL).

    Ap_or_none : host matrix already permuted (A[p][:, p]) on every rank, or None with
                 generator = (seed, mu) to use the device generator of family F2 permuted by p.
    Returns (steps, d, omega_pos, delta_pos, clamped): `steps.S` holds this rank's columns of L.
    omega / delta are returned in POSITION order here (scatter with p for the reference's original order).
    """
    import torch
    import torch.distributed as dist
    from . import _native
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.cuda.current_device()
    main, bulk = _streams(device)
    steps = CudaSteps(n, rank, world, device)
    if Ap_or_none is not None:
        steps.set_matrix(Ap_or_none, main)
    else:
        steps.generate_f2(generator[0], generator[1], p, main)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, bounds_kwargs.get('min_diag_D'), bounds_kwargs.get('max_diag_D'),
                                       bounds_kwargs['min_abs_value_D'], bounds_kwargs['min_abs_value_L'],
                                       bounds_kwargs.get('min_diag_B'), bounds_kwargs.get('max_diag_B'))
    steps.set_bounds(bounds, p, main)
    torch.cuda.synchronize()
    factorize(n, rank, world, steps, NcclComm(), main, bulk, lookahead=lookahead)
    torch.cuda.synchronize()
    d, omega, delta, clamped = steps.finalize(main)
    return steps, d, omega, delta, clamped


class DistributedFactor:
    """The result of `approximate_distributed` as something a caller can USE:
    L stays block-cyclic on the GPUs, `solve` runs there (decompositions.py:983-991), `gather` turns it into the
    reference's LDL_Decomposition on one rank when it fits (Reimer.py:810-812).

    `d`, `p` are replicated numpy arrays (position order / position -> original index); `omega`, `delta` are in
    POSITION order as approximate_distributed returns them.  Collective methods must be called by every rank."""

    SOLVE_WINDOW = 8 * NB

    def __init__(self, steps, d, p, omega=None, delta=None, clamped=None):
        self.steps, self.n = steps, steps.n
        self.d = np.asarray(d, dtype=np.float64)
        self.p = np.arange(self.n, dtype=np.int64) if p is None else np.asarray(p, dtype=np.int64)
        self.omega, self.delta, self.clamped = omega, delta, clamped
        self._solve_ws = None
        self._stream = None

    def close(self):
        self._solve_ws = None
        self.steps.close()

    def _main(self):
        import torch
        if self._stream is None:
            self._stream = TorchStream(torch.cuda.Stream(device=self.steps.device, priority=-1))
        return self._stream

    @staticmethod
    def _all_reduce(t, world):
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t)

    # ---- gather (collective) ----
    def gather(self, dst=0):
        """The whole factor on rank `dst` as an LDL_Decomposition (with the factor resident on its GPU; `.L`
        downloads it), None on the other ranks.  Needs n^2 * 8 bytes plus one slab of free memory on that GPU."""
        import torch
        import torch.distributed as dist
        from . import decompositions
        st = self.steps
        n, world, rank = self.n, st.world, st.rank
        main = self._main()
        lib, ctx = st.ctx.lib, st.ctx
        dec = None
        with torch.cuda.stream(main.stream):
            st._on(main)
            full = torch.zeros((n, n), dtype=torch.float64, device=st.S.device) if rank == dst else None
            for r in range(world):
                ld_r = local_cols(n, r, world)
                if rank == dst:
                    slab = st.S if r == rank else torch.empty((n, ld_r), dtype=torch.float64, device=st.S.device)
                    if r != rank:
                        dist.recv(slab, src=r)
                    ctx.check(lib.mdb_dist_place_columns(ctx.handle, n, r, world, ctypes.c_void_p(slab.data_ptr()), ld_r,
                                                         ctypes.c_void_p(full.data_ptr()), n), 'mdb_dist_place_columns')
                    torch.cuda.current_stream().synchronize()
                    del slab
                elif rank == r:
                    dist.send(st.S, dst=dst)
            if rank == dst:
                h = ctypes.c_void_p()
                P = st._native.ptr
                d64 = np.ascontiguousarray(self.d)
                p64 = np.ascontiguousarray(self.p)
                ctx.check(lib.mdb_factor_upload_device(ctx.handle, n, st._native.TYPE_LDL, ctypes.c_void_p(full.data_ptr()), n,
                                                       P(d64), P(p64), ctypes.byref(h)), 'mdb_factor_upload_device')
                del full
                dec = decompositions.LDL_Decomposition(None, self.d.astype(np.longdouble), p=self.p)
                dec._attach_lazy_matrix(st._native.Factor(ctx, h), n, np.ones(n))
                if self.omega is not None:      # original index order like R.py:538,545
                    for name in ('omega', 'delta'):
                        v = np.empty(n, dtype=np.longdouble)
                        v[self.p] = getattr(self, name)
                        setattr(dec, name, v)
                    dec.clamped = self.clamped
        return dec

    # ---- distributed solve (collective) ----
    def _prepare_solve(self):
        import torch
        if self._solve_ws is not None:
            return self._solve_ws
        st = self.steps
        lib, ctx, h = st.ctx.lib, st.ctx, st.handle
        dev = st.S.device
        main = self._main()
        ws = dict(windows=int(lib.mdb_dist_solve_windows(h)))
        with torch.cuda.stream(main.stream):
            st._on(main)
            ws['band'] = torch.empty(int(lib.mdb_dist_solve_band_elems(h)), dtype=torch.float64, device=dev)
            ws['tinv'] = torch.empty(int(lib.mdb_dist_solve_tinv_elems(h)), dtype=torch.float64, device=dev)
            ws['partial'] = torch.empty(int(lib.mdb_dist_solve_partial_elems(h)), dtype=torch.float64, device=dev)
            ctx.check(lib.mdb_dist_solve_fill_band(h, ctypes.c_void_p(ws['band'].data_ptr())), 'mdb_dist_solve_fill_band')
            self._all_reduce(ws['band'], st.world)     # disjoint supports: the sum is exact
            ctx.check(lib.mdb_dist_solve_invert_band(h, ctypes.c_void_p(ws['band'].data_ptr()),
                                                     ctypes.c_void_p(ws['tinv'].data_ptr())), 'mdb_dist_solve_invert_band')
        self._solve_ws = ws
        return ws

    def solve(self, b):
        """x with A x = b for the factorized (approximated) matrix, b of shape (n,) or (n, k), the same on every rank;
        returns x on every rank.  Same contract as DecompositionBase.solve (decompositions.py:738-766)."""
        import torch
        from . import errors
        st = self.steps
        n, world = self.n, st.world
        b = np.asarray(b)
        B = np.ascontiguousarray(b.reshape(n, -1), dtype=np.float64)
        if np.any(np.abs(self.d) < np.finfo(np.longdouble).resolution):       # decompositions.py:949-951
            raise errors.DecompositionSingularError(self)
        ws = self._prepare_solve()
        lib, ctx, h = st.ctx.lib, st.ctx, st.handle
        dev = st.S.device
        main = self._main()
        WD = self.SOLVE_WINDOW
        X = np.empty_like(B)
        q = np.empty(n, dtype=np.int64)
        q[self.p] = np.arange(n)
        vp = ctypes.c_void_p
        with torch.cuda.stream(main.stream):
            st._on(main)
            for c0 in range(0, B.shape[1], 8):
                kc = min(8, B.shape[1] - c0)
                k = 1 if kc == 1 else 2 if kc == 2 else 4 if kc <= 4 else 8
                chunk = np.zeros((n, k))
                chunk[:, :kc] = B[self.p, c0:c0 + kc]                      # b[p]
                r = torch.from_numpy(chunk).to(dev)
                t = torch.zeros((n, k), dtype=torch.float64, device=dev)
                xchg = torch.empty((WD, k), dtype=torch.float64, device=dev)
                for w in range(ws['windows']):
                    self._all_reduce(t[w * WD:min((w + 1) * WD, n)], world)
                    ctx.check(lib.mdb_dist_solve_fwd_window(h, w, vp(ws['band'].data_ptr()), vp(ws['tinv'].data_ptr()),
                                                            vp(r.data_ptr()), vp(t.data_ptr()), k), 'mdb_dist_solve_fwd_window')
                ctx.check(lib.mdb_dist_solve_scale(h, vp(r.data_ptr()), k), 'mdb_dist_solve_scale')
                for w in range(ws['windows'] - 1, -1, -1):
                    ctx.check(lib.mdb_dist_solve_bwd_colsum(h, w, vp(r.data_ptr()), vp(xchg.data_ptr()),
                                                            vp(ws['partial'].data_ptr()), k), 'mdb_dist_solve_bwd_colsum')
                    self._all_reduce(xchg, world)
                    ctx.check(lib.mdb_dist_solve_bwd_window(h, w, vp(ws['band'].data_ptr()), vp(ws['tinv'].data_ptr()),
                                                            vp(r.data_ptr()), vp(xchg.data_ptr()), k), 'mdb_dist_solve_bwd_window')
                X[:, c0:c0 + kc] = r.cpu().numpy()[q, :kc]                 # x[p^-1]
        return X.reshape(b.shape)


def decomposition_distributed(A, n=None, generator=None, lookahead=True, **bounds_kwargs):
    """`approximation.positive_semidefinite.decomposition` over all ranks of the default process group, for matrices
    that need more than one GPU (BASELINE config 4): returns a DistributedFactor (solve on the slabs, gather when it
    fits).  `A`: the host matrix, the same on every rank (or None with generator = (seed, mu) for family F2);
    static permutation methods / explicit vectors only (dynamic pivoting would need an arg-min all-reduce per column:
    replicas only)."""
    from . import permute
    from .approximation.positive_semidefinite import reimer
    permutation = bounds_kwargs.pop('permutation', 'decreasing_diagonal_values')
    if A is not None:
        A = np.asarray(A)
        n = A.shape[0]
        prepared = reimer._prepare(A, bounds_kwargs.get('min_diag_B'), bounds_kwargs.get('max_diag_B'),
                                   bounds_kwargs.get('min_diag_D'), bounds_kwargs.get('max_diag_D'),
                                   bounds_kwargs.get('min_abs_value_D'), bounds_kwargs.get('min_abs_value_L'), permutation)
        _, _, p, dynamic, mB, MB, mD, MD, mavD, mavL = prepared
        if dynamic:
            raise ValueError('the dynamic pivoting methods are not available for the distributed factorization; pass a '
                             'static permutation method or a permutation vector.')
        p = np.asarray(p, dtype=np.int64)
        Ap = np.ascontiguousarray(A[p[:, None], p[None, :]], dtype=np.float64)
        kw = dict(min_diag_D=mD, max_diag_D=MD, min_abs_value_D=mavD, min_abs_value_L=mavL,
                  min_diag_B=None if np.all(np.isneginf(mB)) else mB, max_diag_B=None if np.all(np.isposinf(MB)) else MB)
    else:
        from . import _native
        seed, mu = generator
        ctx = _native.context()
        diag = np.empty(n)
        ctx.check(ctx.lib.mdb_generate_diag_f2(ctx.handle, n, seed, mu, _native.ptr(diag)), 'mdb_generate_diag_f2')
        p = np.asarray(permute.permutation_vector(_DiagOnly(diag), permutation) if isinstance(permutation, str)
                       else permutation, dtype=np.int64)
        Ap = None
        kw = dict(bounds_kwargs)
        kw.setdefault('min_abs_value_D', float(np.finfo(np.longdouble).eps) ** 0.5)
        kw.setdefault('min_abs_value_L', float(np.finfo(np.float64).eps))
    steps, d, omega, delta, clamped = approximate_distributed(Ap, n, kw, p=p, generator=generator, lookahead=lookahead)
    return DistributedFactor(steps, d, p, omega, delta, clamped)


class _DiagOnly:
    """Just enough of a matrix for permute.permutation_vector: shape and diagonal()."""

    def __init__(self, diag):
        self._diag = diag
        self.shape = (len(diag), len(diag))

    def diagonal(self):
        return self._diag


def _apply_composed_from_A(steps, omega, delta, V):
    """(A o Omega + diag(gamma + delta)) V and A V in position order from the local column slabs of the ORIGINAL
    matrix (steps.A), one all-reduce each -- the right-hand side of Reimer.py:947-957.  Test / bench helper (plain
    torch products on the slabs), not part of the factorization or solve path."""
    import torch
    import torch.distributed as dist
    n, ld, rank, world = steps.n, steps.ld, steps.rank, steps.world
    dev = steps.S.device
    nprobe = V.shape[1]
    lc = torch.arange(ld, device=dev)
    gc = ((lc // NB) * world + rank) * NB + lc % NB
    valid = gc < n
    gcv = gc.clamp(max=n - 1)
    w_t = torch.as_tensor(omega, device=dev)
    de_t = torch.as_tensor(delta, device=dev)
    # (A o Omega)[i, k] = A[i, k] * omega[max(i, k)], diagonal replaced by gamma + delta
    rows = torch.arange(n, device=dev)
    rhs = torch.zeros((n, nprobe), dtype=torch.float64, device=dev)
    av = torch.zeros((n, nprobe), dtype=torch.float64, device=dev)
    chunk = 4096
    for c0 in range(0, ld, chunk):
        c1 = min(c0 + chunk, ld)
        Ablk = steps.A[:, c0:c1]
        g_blk = gcv[c0:c1]
        W = torch.maximum(rows[:, None], g_blk[None, :])
        Om = w_t[W]
        diag_mask = rows[:, None] == g_blk[None, :]
        M = torch.where(diag_mask, (Ablk + de_t[g_blk][None, :]), Ablk * Om) * valid[c0:c1][None, :]
        rhs += M @ V[g_blk] * 1.0
        av += (Ablk * valid[c0:c1][None, :]) @ V[g_blk]
        del W, Om, M
    if world > 1:
        dist.all_reduce(rhs)
        dist.all_reduce(av)
    return rhs, av, gcv, valid


def identity_residual(steps, d, omega, delta, nprobe=2, seed=5):
    """Oracle-free check at any n (Reimer.py:947-957): L D L^T v == (A o Omega + diag(gamma + delta)) v in
    position order, with L and A as local column slabs and one all-reduce per side.  Test / bench helper
    (plain torch products on the slabs), not part of the factorization path."""
    import torch
    import torch.distributed as dist
    n = steps.n
    dev = steps.S.device
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    V = torch.rand((n, nprobe), generator=g, dtype=torch.float64).to(dev) * 2 - 1
    rhs, av, gcv, valid = _apply_composed_from_A(steps, omega, delta, V)
    d_t = torch.as_tensor(d, device=dev)
    L = steps.S                                   # local columns of L after finalize
    z = (L.T @ V) * (d_t[gcv] * valid)[:, None]   # d_k (L[:, k] . v) for local k
    lhs = L @ z
    if steps.world > 1:
        dist.all_reduce(lhs)
    return float((lhs - rhs).abs().max() / av.abs().max())


def solve_residual(factor, b_pos, x_pos):
    """|| (A o Omega + diag(gamma + delta)) x - b ||_max / || b ||_max in position order: the distributed solve checked
    against the ORIGINAL matrix slabs through the identity B = A o Omega + diag(delta) (Reimer.py:947-957), so it
    does not use L at all.  Test / bench helper."""
    import torch
    dev = factor.steps.S.device
    X = torch.as_tensor(np.ascontiguousarray(x_pos.reshape(factor.n, -1)), device=dev)
    Bx, _, _, _ = _apply_composed_from_A(factor.steps, factor.omega, factor.delta, X)
    B = torch.as_tensor(np.ascontiguousarray(b_pos.reshape(factor.n, -1)), device=dev)
    return float((Bx - B).abs().max() / B.abs().max())


def _mem_available_bytes():
    try:
        for line in open('/proc/meminfo'):
            if line.startswith('MemAvailable:'):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def bench_f2(n, s, steps, warmup, seed, e2e=True, peak_tflops=None, peak_hbm_gbs=None):
    """bench.py's N > 1 leg: family F2 generated on the devices, timed with CUDA events, max over ranks.
    Adds a serialised profiling pass (roofline of the bulk update kernel on this rank's share) and an
    end-to-end leg in which every rank starts from its own column slab in pinned HOST memory."""
    import torch
    import torch.distributed as dist
    from . import _native
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.cuda.current_device()
    dev = torch.device('cuda', device)
    ctx = _native.context(device)
    mu = s * math.sqrt(2 * n)
    diag = np.empty(n)
    ctx.check(ctx.lib.mdb_generate_diag_f2(ctx.handle, n, seed, mu, _native.ptr(diag)), 'mdb_generate_diag_f2')
    p = np.argsort(-diag).astype(np.int64)
    main, bulk = _streams(device)
    st = CudaSteps(n, rank, world, device)
    bounds, keep = _native.make_bounds(_native.RULE_REIMER, 0.015 * mu, 0.9 * mu, 0.015 * mu,
                                       float(np.finfo(np.float64).eps), None, 2 * mu)
    comm = NcclComm()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    times = []
    launches = 0
    for it in range(warmup + steps):
        st.generate_f2(seed, mu, p, main)
        st.set_bounds(bounds, p, main)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main.stream)
        factorize(n, rank, world, st, comm, main, bulk, lookahead=True)
        e1.record(main.stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = max_over_ranks(e0.elapsed_time(e1))
        if it >= warmup:
            times.append(t)
            launches += ctx.launch_count() - l0
    d, omega, delta, clamped = st.finalize(main)
    ident = identity_residual(st, d, omega, delta)
    # ---- the consumer of the distributed factor: solve on the slabs (1 right-hand side), checked against the
    #      original matrix through B = A o Omega + diag(delta) ----
    dsolve = None
    try:
        import time as _time
        hbm = peak_hbm_gbs or 6547.5
        fac = DistributedFactor(st, d, np.arange(n, dtype=np.int64), omega, delta, clamped)   # position order in and out
        b = np.random.default_rng(1).standard_normal(n)
        x = fac.solve(b)                       # builds the replicated band (once per factor)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = _time.perf_counter()
        x = fac.solve(b)
        torch.cuda.synchronize()
        sec = max_over_ranks(_time.perf_counter() - t0)
        gbs = 8.0 * n * n / sec / 1e9
        dsolve = dict(n=n, nrhs=1, seconds=sec, algorithmic_bytes=8.0 * n * n, achieved=gbs, peak=hbm * world, unit='GB/s',
                      frac=gbs / (hbm * world), residual_vs_original_matrix=solve_residual(fac, b, x),
                      note='host-timed (max over ranks), b and x as numpy arrays; L streamed once per sweep from the '
                           'block-cyclic slabs, one 1024-row window = one small all-reduce per sweep')
        fac._solve_ws = None
        del fac
    except Exception as e:  # noqa: BLE001
        dsolve = dict(error='%s: %s' % (type(e).__name__, e))
    out = dict(ms_per_step=float(np.mean(times)), gpu_launches=int(launches), identity=ident,
               n_clamped=int(clamped.sum()), dist_solve=dsolve,
               panel_exchange='NCCL broadcast')

    # ---- roofline: one serialised pass (no lookahead, everything on the main stream), events per bulk update ----
    st.generate_f2(seed, mu, p, main)
    st.set_bounds(bounds, p, main)
    torch.cuda.synchronize()
    dist.barrier()
    st.bulk_events = []
    factorize(n, rank, world, st, comm, main, bulk, lookahead=False)
    torch.cuda.synchronize()
    bulk_ms = sum(a.elapsed_time(b) for a, b in st.bulk_events)
    n_bulk = len(st.bulk_events)
    st.bulk_events = None
    stats = ctx.update_stats()
    dist.barrier()
    if bulk_ms > 0 and n_bulk:
        achieved = stats['bulk_flop'] / (bulk_ms * 1e-3) / 1e12
        out['roofline'] = dict(bound='tensor', kernel='k_gemm_tn_tma (bulk trailing update on rank 0\'s column blocks)',
                               achieved=achieved, peak=peak_tflops, unit='TFLOP/s',
                               frac=(achieved / peak_tflops if peak_tflops else None), launches=n_bulk,
                               avg_launch_ms=bulk_ms / n_bulk,
                               algorithmic_flop_per_launch='2 K x (strictly-lower elements of this rank\'s column blocks '
                                                           'beyond the outer block), K = 128 x its panels',
                               algorithmic_flop_total=stats['bulk_flop'],
                               share_of_n3_over_3=stats['bulk_flop'] * world / (n ** 3 / 3), traffic=None,
                               note='per rank (rank 0); serialised pass without lookahead')

    # ---- end to end: every rank starts from its own column slab in pinned host memory ----
    out['e2e'] = None
    if e2e:
        slab_bytes = n * st.ld * 8
        local_ranks = int(os.environ.get('LOCAL_WORLD_SIZE', world))
        ok = _mem_available_bytes() > 1.5 * slab_bytes * local_ranks + (16 << 30)
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() > 0:
            host = ctx.pinned_array((n, st.ld))
            st.generate_f2(seed, mu, p, main)            # fill the host slab once, untimed
            torch.cuda.synchronize()
            ctx.check(ctx.lib.mdb_memcpy(ctx.handle, _native.ptr(host), ctypes.c_void_p(st.A.data_ptr()), slab_bytes,
                                         _native.HOST, _native.DEVICE), 'mdb_memcpy')
            e2e_times = []
            import time
            for it in range(2):                          # one warm-up call, one timed
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                st.set_local(host, main, lower_only=True)   # H2D of the rows the factorization reads (+ device copy)
                st.set_bounds(bounds, p, main)
                factorize(n, rank, world, st, comm, main, bulk, lookahead=True)
                dv, ov, de, cl = st.finalize(main)       # D2H of d, omega, delta, clamped; L stays distributed on the GPUs
                torch.cuda.synchronize()
                dt = max_over_ranks(time.perf_counter() - t0)
                if it > 0:
                    e2e_times.append(dt)
            _native.free_pinned(host)
            sec = float(np.mean(e2e_times))
            out['e2e'] = dict(value=n ** 3 / 3 / sec / 1e9, unit='GFLOP/s',
                              h2d_bytes_per_step=int(sum((n - g * NB) * min(NB, n - g * NB) * 8 for g in range((n + NB - 1) // NB))),
                              h2d_note='per column block only the rows from its diagonal block downwards cross PCIe (the host '
                                       'slabs hold n x local columns = {} bytes in total)'.format(
                                           int(sum(n * local_cols(n, r, world) * 8 for r in range(world)))),
                              d2h_bytes_per_step=int(world * (3 * n * 8 + n)), seconds_per_step=sec, steps=1,
                              call='distributed.CudaSteps.set_local (pinned host column slabs) + factorize + finalize; '
                                   'L stays distributed on the devices (SURVEY 8e), the vectors d/omega/delta/clamped '
                                   'return to every host', n_clamped=int(cl.sum()), results_finite=bool(np.all(np.isfinite(dv))))
        else:
            out['e2e_skipped'] = ('host memory: %.0f GB available for %d x %.0f GB pinned slabs'
                                  % (_mem_available_bytes() / 1e9, local_ranks, slab_bytes / 1e9))
    st.close()
    return out
