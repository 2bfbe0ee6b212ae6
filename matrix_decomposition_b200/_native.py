"""ctypes binding of libmdb200.so (include/mdb200.h).

The library is the ONLY compute path of this package: if it is missing, or no CUDA device is
usable, every operation raises `errors.NativeLibraryError`.  There is deliberately no numpy/CPU
fallback (the oracle under oracle/ is test infrastructure and is never imported from here).
"""
import ctypes
import math
import os
import threading

import numpy as np

from . import errors

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmdb200.so')

MDB_OK = 0
MDB_ERR_INVALID = 1
MDB_ERR_CUDA = 2
MDB_ERR_NO_DEVICE = 3
MDB_ERR_NOT_POSITIVE_DEFINITE = 4
MDB_ERR_STATE = 5

RULE_REIMER = 0
RULE_EXACT = 1
TYPE_LDL = 0
TYPE_LDL_COMPRESSED = 1
TYPE_LL = 2
TYPE_NONE = -1
HOST = 0
DEVICE = 1

TYPE_CODES = {'LDL': TYPE_LDL, 'LDL_compressed': TYPE_LDL_COMPRESSED, 'LL': TYPE_LL}


class Bounds(ctypes.Structure):
    """struct mdb_bounds"""
    _fields_ = [('rule', ctypes.c_int32),
                ('min_diag_D', ctypes.c_double),
                ('max_diag_D', ctypes.c_double),
                ('min_abs_value_D', ctypes.c_double),
                ('min_abs_value_L', ctypes.c_double),
                ('min_diag_B', ctypes.c_void_p),
                ('max_diag_B', ctypes.c_void_p),
                ('stride_B', ctypes.c_int64)]


_i64, _dbl, _vp, _int = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
_u64, _sz = ctypes.c_uint64, ctypes.c_size_t
_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); must list every symbol that include/mdb200.h declares
SIGNATURES = {
    'mdb_version': (_int, []),
    'mdb_ctx_create': (_int, [_int, _pp]),
    'mdb_ctx_destroy': (_int, [_vp]),
    'mdb_ctx_set_stream': (_int, [_vp, _vp]),
    'mdb_ctx_trim': (_int, [_vp]),
    'mdb_ctx_synchronize': (_int, [_vp]),
    'mdb_last_error': (ctypes.c_char_p, [_vp]),
    'mdb_launch_count': (_i64, [_vp]),
    'mdb_ctx_set_profiling': (_int, [_vp, _int]),
    'mdb_ctx_get_phase_ms': (_int, [_vp, _vp]),
    'mdb_ctx_set_blocking': (_int, [_vp, _int, _i64, _i64, _i64]),
    'mdb_ctx_get_update_stats': (_int, [_vp, _vp]),
    'mdb_ctx_last_input_nonfinite': (_int, [_vp]),
    'mdb_malloc': (_int, [_vp, _sz, _pp]),
    'mdb_free': (_int, [_vp, _vp]),
    'mdb_malloc_host': (_int, [_vp, _sz, _pp]),
    'mdb_free_host': (_int, [_vp, _vp]),
    'mdb_memcpy': (_int, [_vp, _vp, _vp, _sz, _int, _int]),
    'mdb_memcpy2d': (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _int, _int]),
    'mdb_permute_sym_f64': (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _int]),
    'mdb_approximate_f64': (_int, [_vp, _i64, _vp, _i64, _vp, ctypes.POINTER(Bounds), _int, _vp, _i64, _vp, _vp, _vp,
                                   _vp, _int, _pp]),
    'mdb_approximate_dynamic_f64': (_int, [_vp, _i64, _vp, _i64, _int, ctypes.POINTER(Bounds), _int, _vp, _i64, _vp, _vp, _vp,
                                           _vp, _vp, _int, _pp]),
    'mdb_decompose_f64': (_int, [_vp, _i64, _vp, _i64, _vp, _int, _vp, _i64, _vp, _vp, _int, _pp]),
    'mdb_approximate_batched_f64': (_int, [_vp, _i64, _i64, _vp, _i64, _vp, ctypes.POINTER(Bounds), _vp, _i64, _vp,
                                           _vp, _vp, _vp, _int]),
    'mdb_factor_create': (_int, [_vp, _i64, _i64, _pp]),
    'mdb_factor_destroy': (_int, [_vp]),
    'mdb_factor_set_matrix': (_int, [_vp, _vp, _i64, _vp, _int]),
    'mdb_generate_diag_f2': (_int, [_vp, _i64, _u64, _dbl, _vp]),
    'mdb_factor_generate_f2': (_int, [_vp, _u64, _dbl, _vp]),
    'mdb_generate_f2': (_int, [_vp, _i64, _u64, _dbl, _vp, _i64]),
    'mdb_factor_run': (_int, [_vp, ctypes.POINTER(Bounds)]),
    'mdb_gmw_f64': (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _int, _dbl, _dbl, _vp, _vp]),
    'mdb_se_f64': (_int, [_vp, _i64, _vp, _i64, _int, _int, _int, _int, _dbl, _dbl, _vp, _vp]),
    'mdb_factor_psd_matrix': (_int, [_vp, _vp, _vp, _i64, _vp, _i64, _int]),
    'mdb_factor_get': (_int, [_vp, _int, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    'mdb_factor_upload': (_int, [_vp, _i64, _int, _vp, _i64, _vp, _vp, _pp]),
    'mdb_factor_solve': (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _int]),
    'mdb_factor_multiply': (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _int]),
    'mdb_factor_composed': (_int, [_vp, _vp, _i64, _int]),
    'mdb_factor_identity_residual': (_int, [_vp, _int, _u64, _vp]),
    'mdb_factor_device_ptr': (_vp, [_vp, _int]),
    'mdb_factor_ld': (_i64, [_vp]),
    'mdb_debug_gemm_tn': (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _int, _int, _int, _i64, _i64,
                                 _int, _vp]),
    'mdb_dist_create': (_int, [_vp, _i64, _int, _int, _vp, _vp, _i64, _vp, _vp, _pp]),
    'mdb_dist_destroy': (_int, [_vp]),
    'mdb_dist_local_cols': (_i64, [_i64, _int, _int]),
    'mdb_dist_panel_elems': (_i64, [_vp, _i64]),
    'mdb_dist_generate_f2': (_int, [_vp, _u64, _dbl, _vp]),
    'mdb_dist_set_matrix': (_int, [_vp, _vp, _i64]),
    'mdb_dist_set_local': (_int, [_vp, _vp, _i64, _int]),
    'mdb_dist_set_local_lower': (_int, [_vp, _vp, _i64, _int]),
    'mdb_dist_set_bounds': (_int, [_vp, ctypes.POINTER(Bounds), _vp]),
    'mdb_dist_outer_width': (_i64, [_vp, _i64]),
    'mdb_dist_factor_panel': (_int, [_vp, _i64, _int, _i64, _int]),
    'mdb_dist_absorb_panel': (_int, [_vp, _i64, _int, _i64, _int]),
    'mdb_dist_apply_update': (_int, [_vp, _i64, _int, _i64, _i64, _i64, _i64]),
    'mdb_dist_finalize': (_int, [_vp, _vp, _vp, _vp, _vp]),
    'mdb_dist_info': (_int, [_vp, _vp]),
    'mdb_dist_place_columns': (_int, [_vp, _i64, _int, _int, _vp, _i64, _vp, _i64]),
    'mdb_factor_upload_device': (_int, [_vp, _i64, _int, _vp, _i64, _vp, _vp, _pp]),
    'mdb_dist_solve_band_elems': (_i64, [_vp]),
    'mdb_dist_solve_tinv_elems': (_i64, [_vp]),
    'mdb_dist_solve_partial_elems': (_i64, [_vp]),
    'mdb_dist_solve_windows': (_i64, [_vp]),
    'mdb_dist_solve_fill_band': (_int, [_vp, _vp]),
    'mdb_dist_solve_invert_band': (_int, [_vp, _vp, _vp]),
    'mdb_dist_solve_fwd_window': (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _int]),
    'mdb_dist_solve_scale': (_int, [_vp, _vp, _int]),
    'mdb_dist_solve_bwd_colsum': (_int, [_vp, _i64, _vp, _vp, _vp, _int]),
    'mdb_dist_solve_bwd_window': (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _int]),
}

_lib = None
_lock = threading.Lock()
_contexts = {}


def load_library():
    """Loads libmdb200.so and declares every prototype.  Raises NativeLibraryError when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise errors.NativeLibraryError(
                'libmdb200.so not found at {} -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a).  This package has no CPU fallback.'.format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def ptr(a):
    """Device-agnostic pointer of a numpy array (or None)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)


class Context:
    """One mdb_ctx per (process, device)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.mdb_ctx_create(int(device), ctypes.byref(h))
        self.handle = h
        self.device = int(device)
        if rc != MDB_OK:
            msg = self.lib.mdb_last_error(h).decode() if h else ''
            if h:
                self.lib.mdb_ctx_destroy(h)
                self.handle = None
            raise errors.NativeLibraryError(
                'no usable CUDA device (mdb_ctx_create(device={}) returned {}{}). This package runs on the GPU only; '
                'there is no CPU fallback.'.format(device, rc, ': ' + msg if msg else ''))

    def check(self, rc, what=''):
        if rc != MDB_OK:
            msg = self.lib.mdb_last_error(self.handle).decode()
            if rc == MDB_ERR_INVALID:
                raise ValueError('{}: {}'.format(what, msg))
            raise errors.NativeLibraryError('{} failed with status {}: {}'.format(what, rc, msg))

    def last_error(self):
        return self.lib.mdb_last_error(self.handle).decode()

    def launch_count(self):
        return int(self.lib.mdb_launch_count(self.handle))

    def synchronize(self):
        self.check(self.lib.mdb_ctx_synchronize(self.handle), 'mdb_ctx_synchronize')

    def trim(self):
        """Returns the cached working buffers (mdb_ctx_trim) to the driver."""
        self.check(self.lib.mdb_ctx_trim(self.handle), 'mdb_ctx_trim')

    def set_stream(self, cuda_stream):
        self.check(self.lib.mdb_ctx_set_stream(self.handle, ctypes.c_void_p(cuda_stream or 0)), 'mdb_ctx_set_stream')

    def set_profiling(self, on):
        self.check(self.lib.mdb_ctx_set_profiling(self.handle, int(bool(on))), 'mdb_ctx_set_profiling')

    def set_blocking(self, fixed=0, t2=4096, t4=8192, t8=32768):
        """Outer-block width of the factorization (see mdb_ctx_set_blocking in include/mdb200.h)."""
        self.check(self.lib.mdb_ctx_set_blocking(self.handle, int(fixed), int(t2), int(t4), int(t8)), 'mdb_ctx_set_blocking')

    def update_stats(self):
        out = (ctypes.c_double * 4)()
        self.check(self.lib.mdb_ctx_get_update_stats(self.handle, out), 'mdb_ctx_get_update_stats')
        return dict(bulk_flop=out[0], bulk_launches=int(out[1]), inner_flop=out[2], inner_launches=int(out[3]))

    def phase_ms(self):
        out = (ctypes.c_double * 4)()
        self.check(self.lib.mdb_ctx_get_phase_ms(self.handle, out), 'mdb_ctx_get_phase_ms')
        return dict(diag=out[0], panel=out[1], trailing=out[2], inner=out[3])

    # --- raw device memory (used by bench.py and the distributed driver) ---
    def malloc(self, nbytes):
        p = ctypes.c_void_p()
        self.check(self.lib.mdb_malloc(self.handle, int(nbytes), ctypes.byref(p)), 'mdb_malloc')
        return p.value

    def free(self, dptr):
        self.check(self.lib.mdb_free(self.handle, ctypes.c_void_p(dptr)), 'mdb_free')

    def malloc_host(self, nbytes):
        p = ctypes.c_void_p()
        self.check(self.lib.mdb_malloc_host(self.handle, int(nbytes), ctypes.byref(p)), 'mdb_malloc_host')
        return p.value

    def free_host(self, hptr):
        self.check(self.lib.mdb_free_host(self.handle, ctypes.c_void_p(hptr)), 'mdb_free_host')

    def pinned_array(self, shape, dtype=np.float64):
        """numpy array backed by pinned host memory (kept alive by the returned array's base)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if len(shape) else 1
        nbytes = max(n * dtype.itemsize, 8)
        hptr = self.malloc_host(nbytes)
        buf = (ctypes.c_char * nbytes).from_address(hptr)
        arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        _PinnedOwner(self, hptr, buf)  # registers itself
        return arr


class _PinnedOwner:
    _live = {}

    def __init__(self, ctx, hptr, buf):
        self.ctx, self.hptr = ctx, hptr
        _PinnedOwner._live[hptr] = self


def free_pinned(arr):
    hptr = arr.ctypes.data
    owner = _PinnedOwner._live.pop(hptr, None)
    if owner is not None:
        owner.ctx.free_host(hptr)


def context(device=None):
    """The process-wide context of `device` (default: $MDB200_DEVICE, else LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get('MDB200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = Context(device)
            _contexts[device] = ctx
        return ctx


def make_bounds(rule, min_diag_D, max_diag_D, min_abs_value_D, min_abs_value_L, min_diag_B=None, max_diag_B=None):
    """Builds a struct mdb_bounds; returns (struct, keepalive)."""
    b = Bounds()
    b.rule = rule
    b.min_diag_D = math.nan if min_diag_D is None else float(min_diag_D)
    b.max_diag_D = math.inf if max_diag_D is None else float(max_diag_D)
    b.min_abs_value_D = float(min_abs_value_D)
    b.min_abs_value_L = float(min_abs_value_L)
    mB = np.ascontiguousarray(np.asarray(-math.inf if min_diag_B is None else min_diag_B, dtype=np.float64).reshape(-1))
    MB = np.ascontiguousarray(np.asarray(math.inf if max_diag_B is None else max_diag_B, dtype=np.float64).reshape(-1))
    if mB.size != MB.size:  # one scalar, one vector -> broadcast the scalar
        n = max(mB.size, MB.size)
        mB = np.ascontiguousarray(np.broadcast_to(mB, (n,)))
        MB = np.ascontiguousarray(np.broadcast_to(MB, (n,)))
    b.min_diag_B = mB.ctypes.data
    b.max_diag_B = MB.ctypes.data
    b.stride_B = 0 if mB.size == 1 else 1
    return b, (mB, MB)


class Factor:
    """Owner of a device-resident mdb_factor handle."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.handle = handle

    def close(self):
        if self.handle:
            self.ctx.lib.mdb_factor_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def download(self, n, type_code):
        """The n x n factor matrix (L or LD, in the type the factor was finalised with) as a fresh numpy array."""
        out = np.empty((n, n), dtype=np.float64)
        rc = self.ctx.lib.mdb_factor_get(self.handle, int(type_code), ptr(out), n, HOST, None, None, None, None, None)
        self.ctx.check(rc, 'mdb_factor_get')
        return out

    def solve(self, b):
        b = np.asarray(b)
        b2 = np.ascontiguousarray(b.reshape(b.shape[0], -1), dtype=np.float64)
        x = np.empty_like(b2)
        rc = self.ctx.lib.mdb_factor_solve(self.handle, ptr(b2), b2.shape[1], b2.shape[1], ptr(x), b2.shape[1], HOST)
        self.ctx.check(rc, 'mdb_factor_solve')
        return x.reshape(b.shape)

    def multiply(self, b):
        b = np.asarray(b)
        b2 = np.ascontiguousarray(b.reshape(b.shape[0], -1), dtype=np.float64)
        x = np.empty_like(b2)
        rc = self.ctx.lib.mdb_factor_multiply(self.handle, ptr(b2), b2.shape[1], b2.shape[1], ptr(x), b2.shape[1], HOST)
        self.ctx.check(rc, 'mdb_factor_multiply')
        return x.reshape(b.shape)

    def composed(self, n):
        B = np.empty((n, n), dtype=np.float64)
        rc = self.ctx.lib.mdb_factor_composed(self.handle, ptr(B), n, HOST)
        self.ctx.check(rc, 'mdb_factor_composed')
        return B
