"""Decomposition objects: LDL_Decomposition, LDL_DecompositionCompressed, LL_Decomposition.

Host-side mirror of the reference's matrix/decompositions.py (same class names, constructor
arguments, properties, methods and error behaviour; file:line citations below refer to that file).
The objects hold numpy arrays exactly like the reference's; the O(n^2)-and-up work -- solve,
inverse-apply, multiply, composed_matrix -- runs on the B200 through libmdb200.so with the factor
kept resident in HBM (uploaded on first use, or handed over by the factorization that produced it).
Dense, real (float64) factors only; sparse and complex factors are outside the hot path.
"""
import abc
import copy as _copy
import os
import tarfile
import tempfile

import numpy as np

from . import _native, constants, errors, permute, util

__all__ = ['DecompositionBase', 'LDL_Decomposition', 'LDL_DecompositionCompressed', 'LL_Decomposition',
           'save', 'load']


def _real_f64_matrix(L, name):
    if np.iscomplexobj(L):
        raise ValueError('{} is complex; this package covers the dense real (float64) path only.'.format(name))
    return L


class DecompositionBase(metaclass=abc.ABCMeta):
    """ A matrix decomposition (decompositions.py:22-787). """

    type_str = constants.BASE_DECOMPOSITION_TYPE

    def __init__(self, p=None):
        self.p = p

    def __str__(self):
        return '{type_str} decomposition of matrix with shape ({n}, {n})'.format(type_str=self.type_str, n=self.n)

    # *** device residency *** #
    # A decomposition produced by this package's factorizations keeps its n x n factor in HBM: solve, multiply and
    # composed_matrix work on it there, and the numpy array behind .L / .LD is only downloaded when it is first
    # read ("lazy").  At n = 65536 that is 34 GB of device -> host copy the reference's users of solve() never see.

    _matrix_attr = None      # '_L' or '_LD'
    _matrix_type_code = None

    def _drop_device_factor(self):
        if '_lazy_matrix' in self.__dict__:      # the only copy of the factor lives on the device: fetch it first
            self._matrix()
        f = self.__dict__.pop('_device_factor', None)
        if f is not None:
            f.close()

    def _attach_device_factor(self, factor):
        """Called by the factorization routines: the factor is already resident in HBM."""
        self._drop_device_factor()
        self.__dict__['_device_factor'] = factor

    def _attach_lazy_matrix(self, factor, n, diagonal):
        """Called by the factorization routines instead of passing the matrix: it stays in HBM until read.
        `diagonal` is the diagonal of the matrix attribute (known without downloading it)."""
        self._drop_device_factor()
        self.__dict__.pop(self._matrix_attr, None)
        self.__dict__['_device_factor'] = factor
        self.__dict__['_lazy_matrix'] = dict(n=int(n), diagonal=np.asarray(diagonal, dtype=np.float64))

    @property
    def is_device_resident_only(self):
        """True while the factor matrix has not been downloaded into a numpy array yet."""
        return '_lazy_matrix' in self.__dict__

    def _matrix(self):
        M = self.__dict__.get(self._matrix_attr)
        if M is None:
            lazy = self.__dict__.get('_lazy_matrix')
            if lazy is None:
                raise AttributeError(self._matrix_attr[1:])
            M = self.__dict__['_device_factor'].download(lazy['n'], self._matrix_type_code)
            self.__dict__[self._matrix_attr] = M
            del self.__dict__['_lazy_matrix']
        return M

    def _matrix_n(self):
        lazy = self.__dict__.get('_lazy_matrix')
        return lazy['n'] if lazy is not None else self._matrix().shape[0]

    def _matrix_diagonal(self):
        lazy = self.__dict__.get('_lazy_matrix')
        return lazy['diagonal'] if lazy is not None else np.asarray(self._matrix()).diagonal()

    @abc.abstractmethod
    def _upload(self, ctx):
        raise NotImplementedError

    def _factor(self):
        f = self.__dict__.get('_device_factor')
        if f is None or not f.handle:
            f = self._upload(_native.context())
            self.__dict__['_device_factor'] = f
        return f

    def __deepcopy__(self, memo):
        if '_lazy_matrix' in self.__dict__:
            self._matrix()
        new = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            if k != '_device_factor':  # device handles are not shared between copies
                new.__dict__[k] = _copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        if '_lazy_matrix' in self.__dict__:
            self._matrix()
        return {k: v for k, v in self.__dict__.items() if k != '_device_factor'}

    # *** permutation (decompositions.py:49-171) *** #

    @property
    def is_permuted(self):
        try:
            p = self._p
        except AttributeError:
            return False
        return bool(np.any(p != np.arange(len(p))))

    @property
    def p(self):
        try:
            return self._p
        except AttributeError:
            return np.arange(self.n)

    @p.setter
    def p(self, p):
        self._drop_device_factor()
        if p is not None:
            self._p = util.as_vector(p)
        else:
            del self.p

    @p.deleter
    def p(self):
        try:
            del self._p
        except AttributeError:
            pass

    @property
    def p_inverse(self):
        return permute.invert_permutation_vector(self.p)

    def _apply_previous_permutation(self, p_previous):
        p_next = self.__dict__.get('_p')
        self._drop_device_factor()
        self._p = permute.concatenate_permutation_vectors(p_previous, p_next)

    def _apply_succeeding_permutation(self, p_next):
        p_previous = self.__dict__.get('_p')
        self._drop_device_factor()
        self._p = permute.concatenate_permutation_vectors(p_previous, p_next)

    @property
    def P(self):
        import scipy.sparse
        p = self.p
        n = len(p)
        P = scipy.sparse.dok_matrix((n, n), dtype=np.int8)
        for i in range(n):
            P[i, p[i]] = 1
        return P

    def permute_matrix(self, A):
        return permute.symmetric(A, self.p) if self.is_permuted else A

    def unpermute_matrix(self, A):
        return permute.symmetric(A, self.p_inverse) if self.is_permuted else A

    # *** basic properties *** #

    @property
    @abc.abstractmethod
    def n(self):
        raise NotImplementedError

    @property
    def composed_matrix(self):
        """The matrix represented by this decomposition (decompositions.py:826-830, :1202-1205), formed on the device
        from the resident factor: the lower triangle of L D L^H block column by block column on the tensor-core tile
        kernel (n^3 / 3 multiply-adds), mirrored, inverse permutation applied (mdb_factor_composed)."""
        return self._factor().composed(self.n)

    # *** compare (decompositions.py:187-230) *** #

    def _is_same_type_and_permutation(self, other):
        return (isinstance(other, DecompositionBase) and self.type_str == other.type_str
                and self.is_permuted == other.is_permuted
                and (not self.is_permuted or (self.n == other.n and bool(np.all(self.p == other.p)))))

    def is_equal(self, other):
        return self._is_same_type_and_permutation(other)

    def __eq__(self, other):
        return self.is_equal(other)

    __hash__ = None

    def is_almost_equal(self, other, rtol=1e-04, atol=1e-06):
        return self._is_same_type_and_permutation(other)

    # *** convert (decompositions.py:232-350) *** #

    def copy(self):
        return _copy.deepcopy(self)

    def is_type(self, type_str):
        return True if type_str is None else self.type_str == type_str

    def as_type(self, type_str, copy=False):
        if self.is_type(type_str):
            return self.copy() if copy else self
        raise errors.NoDecompositionConversionImplementedError(self, type_str)

    def as_any_type(self, *type_strs, copy=False):
        if len(type_strs) == 0 or any(map(self.is_type, type_strs)):
            return self.copy() if copy else self
        return self.as_type(type_strs[0])

    def as_same_type(self, dec, copy=False):
        return dec.as_type(self.type_str, copy=copy)

    # *** features (decompositions.py:352-456) *** #

    def is_sparse(self):
        return False

    @abc.abstractmethod
    def is_positive_semidefinite(self):
        raise NotImplementedError

    def is_positive_definite(self):
        return self.is_positive_semidefinite() and not self.is_singular()

    @abc.abstractmethod
    def is_finite(self):
        raise NotImplementedError

    def check_finite(self, check_finite=True):
        if check_finite and not self.is_finite():
            raise errors.DecompositionNotFiniteError(self)

    @abc.abstractmethod
    def is_singular(self):
        raise NotImplementedError

    def is_invertible(self):
        return not self.is_singular()

    def check_invertible(self):
        if self.is_singular():
            raise errors.DecompositionSingularError(self)

    # *** save and load: uncompressed tar of type.txt + one npz per attribute (decompositions.py:458-615) *** #

    @property
    @abc.abstractmethod
    def _attribute_names(self):
        raise NotImplementedError

    @staticmethod
    def _check_decomposition_filename(filename):
        suffix = '.' + constants.DECOMPOSITION_FILENAME_EXTENSION
        return filename if filename.endswith(suffix) else filename + suffix

    @staticmethod
    def _attribute_filename(attribute_name, is_sparse=False):
        ext = (constants.DECOMPOSITION_ATTRIBUTE_SPARSE_FILE_EXTENSION if is_sparse
               else constants.DECOMPOSITION_ATTRIBUTE_DENSE_FILE_EXTENSION)
        return constants.DECOMPOSITION_ATTRIBUTE_FILENAME.format(attribute_name=attribute_name, file_extension=ext)

    def save(self, filename):
        filename = DecompositionBase._check_decomposition_filename(filename)
        dirname = os.path.dirname(filename)
        if dirname:
            os.makedirs(dirname, exist_ok=True)
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, constants.DECOMPOSITION_TYPE_FILENAME), 'w') as f:
                f.write(self.type_str)
            for name in self._attribute_names:
                np.savez_compressed(os.path.join(tmp, self._attribute_filename(name)), **{name: getattr(self, name)})
            with tarfile.open(filename, mode='w') as tar:
                for file in os.listdir(tmp):
                    tar.add(os.path.join(tmp, file), arcname=file)

    @staticmethod
    def _load_from_tar_file(filename, member, loader):
        try:
            with tarfile.open(filename, mode='r') as tar:
                with tar.extractfile(member) as fobj:
                    return loader(fobj)
        except KeyError as e:
            raise errors.DecompositionInvalidFile(filename) from e

    @staticmethod
    def _load_type(filename):
        filename = DecompositionBase._check_decomposition_filename(filename)
        return DecompositionBase._load_from_tar_file(filename, constants.DECOMPOSITION_TYPE_FILENAME,
                                                     lambda f: f.read().decode())

    def load(self, filename):
        filename = DecompositionBase._check_decomposition_filename(filename)
        type_str = self._load_type(filename)
        if type_str != self.type_str:
            raise errors.DecompositionInvalidDecompositionTypeFile(filename, type_str, self.type_str)
        for name in self._attribute_names:
            def loader(fobj, name=name):
                import io
                with np.load(io.BytesIO(fobj.read()), allow_pickle=False) as z:
                    keys = list(z.keys())
                    if len(keys) != 1 or keys[0] != name:
                        raise errors.DecompositionInvalidFile(filename)
                    return z[name]
            setattr(self, name, self._load_from_tar_file(filename, self._attribute_filename(name), loader))

    # *** multiply / solve (decompositions.py:617-766) *** #

    def _as_real_rhs(self, x, dtype=None):
        x = util.as_matrix_or_vector(x, dtype=dtype, copy=False)
        if np.iscomplexobj(x):
            raise ValueError('complex right-hand sides are outside the dense real (float64) path of this package.')
        if x.shape[0] != self.n:
            raise ValueError('The first dimension of x ({}) must be equal to n ({}).'.format(x.shape[0], self.n))
        return x

    def matrix_right_side_multiplication(self, x, dtype=None):
        """`A @ x` for the composed matrix A (decompositions.py:961-968 LDL, :1324-1330 LL) on the device."""
        x = self._as_real_rhs(x, dtype)
        return self._factor().multiply(x)

    def matrix_both_sides_multiplication(self, x, y=None, dtype=None):
        """`y.H @ A @ x` (decompositions.py:970-981, :1332-1343)."""
        x = self._as_real_rhs(x, dtype)
        y = x if y is None else self._as_real_rhs(y, dtype)
        return util.conjugate_transpose(y) @ self.matrix_right_side_multiplication(x)

    def inverse_matrix_right_side_multiplication(self, x, dtype=None):
        """`A^-1 @ x` (decompositions.py:983-991 LDL, :1345-1352 LL): x[p], forward substitution, / d,
        back substitution, [p^-1] -- on the device, with the factor resident in HBM."""
        x = self._as_real_rhs(x, dtype)
        self.check_invertible()
        return self._factor().solve(x)

    def inverse_matrix_both_sides_multiplication(self, x, y=None, dtype=None):
        """`y.H @ A^-1 @ x` (decompositions.py:993-1005, :1354-1366)."""
        x = self._as_real_rhs(x, dtype)
        y = x if y is None else self._as_real_rhs(y, dtype)
        return util.conjugate_transpose(y) @ self.inverse_matrix_right_side_multiplication(x)

    def solve(self, b, dtype=None):
        """Solves `A x = b` (decompositions.py:738-766).  Raises DecompositionSingularError like the reference."""
        return self.inverse_matrix_right_side_multiplication(b, dtype=dtype)

    @abc.abstractmethod
    def append_block_decomposition(self, dec):
        raise NotImplementedError


class LDL_Decomposition(DecompositionBase):
    """ `L D L^H` of the permuted matrix; L unit lower triangular, d = diag(D)  (decompositions.py:790-1013). """

    type_str = constants.LDL_DECOMPOSITION_TYPE
    _matrix_attr = '_L'
    _matrix_type_code = _native.TYPE_LDL

    def __init__(self, L=None, d=None, p=None):
        self.L = L
        self.d = d
        super().__init__(p=p)

    @property
    def n(self):
        return self._matrix_n()

    @property
    def L(self):
        return self._matrix()

    @L.setter
    def L(self, L):
        self._drop_device_factor()
        if L is not None:
            L = util.as_matrix(_real_f64_matrix(np.asarray(L), 'L'))  # copies, like decompositions.py:842
            if np.any(L.diagonal() != 1):
                raise ValueError('The diagonal values of the lower triangle matrix L must all be equal to 1.')
            self._L = L
        else:
            self.__dict__.pop('_L', None)

    @property
    def d(self):
        return self._d

    @d.setter
    def d(self, d):
        self._drop_device_factor()
        if d is not None:
            d = util.as_vector(d)
            if np.iscomplexobj(d) and np.all(np.isreal(d)):
                d = d.real
            self._d = d
        else:
            self.__dict__.pop('_d', None)

    @property
    def D(self):
        import scipy.sparse
        return scipy.sparse.diags(self.d)

    @property
    def LD(self):
        LD = self.L.copy()
        LD[np.diag_indices(self.n)] = self.d
        return LD

    def is_equal(self, other):
        return (super().is_equal(other) and bool(np.all(self.d == other.d)) and util.is_equal(self.L, other.L))

    def is_almost_equal(self, other, rtol=1e-04, atol=1e-06):
        return (super().is_almost_equal(other) and np.allclose(self.d, other.d, rtol=rtol, atol=atol)
                and util.is_almost_equal(self.L, other.L, rtol=rtol, atol=atol))

    def as_LL_Decomposition(self):
        d, p = self.d, self.p
        for i, d_i in enumerate(d):  # decompositions.py:903-908
            if d_i < 0:
                raise errors.NoDecompositionPossibleWithProblematicSubdecompositionError(
                    self, constants.LL_DECOMPOSITION_TYPE, p[i])
        return LL_Decomposition(self.L * np.sqrt(np.asarray(d, dtype=np.float64))[np.newaxis, :], p=p)

    def as_LDL_DecompositionCompressed(self):
        return LDL_DecompositionCompressed(self.LD, p=self.p)

    def as_type(self, type_str, copy=False):
        try:
            return super().as_type(type_str, copy=copy)
        except errors.NoDecompositionConversionImplementedError:
            if type_str == constants.LL_DECOMPOSITION_TYPE:
                return self.as_LL_Decomposition()
            if type_str == constants.LDL_DECOMPOSITION_COMPRESSED_TYPE:
                return self.as_LDL_DecompositionCompressed()
            raise

    def is_finite(self):
        return util.is_finite(self.L) and util.is_finite(self.d)

    def is_positive_semidefinite(self):
        return bool(np.all(self.d >= 0))

    def is_positive_definite(self):
        return bool(np.all(self.d >= np.finfo(self.d.dtype).resolution))

    def is_singular(self):  # decompositions.py:949-951
        return bool(np.any(np.abs(self.d) < np.finfo(self.d.dtype).resolution))

    @property
    def _attribute_names(self):
        return ('L', 'd', 'p')

    def _upload(self, ctx):
        L = np.ascontiguousarray(self.L, dtype=np.float64)
        d = np.ascontiguousarray(self.d, dtype=np.float64)
        p = np.ascontiguousarray(self.p, dtype=np.int64)
        import ctypes
        h = ctypes.c_void_p()
        rc = ctx.lib.mdb_factor_upload(ctx.handle, self.n, _native.TYPE_LDL, _native.ptr(L), self.n, _native.ptr(d),
                                       _native.ptr(p), ctypes.byref(h))
        ctx.check(rc, 'mdb_factor_upload')
        return _native.Factor(ctx, h)

    def append_block_decomposition(self, dec):
        dec = self.as_same_type(dec, copy=False)
        return LDL_Decomposition(p=np.concatenate((self.p, dec.p + self.n)), L=util.block_diag(self.L, dec.L),
                                 d=np.concatenate((self.d, dec.d)))


class LDL_DecompositionCompressed(DecompositionBase):
    """ `L D L^H` with L and D stored in one matrix LD  (decompositions.py:1016-1167). """

    type_str = constants.LDL_DECOMPOSITION_COMPRESSED_TYPE
    _matrix_attr = '_LD'
    _matrix_type_code = _native.TYPE_LDL_COMPRESSED

    def __init__(self, LD=None, p=None):
        self.LD = LD
        super().__init__(p=p)

    @property
    def n(self):
        return self._matrix_n()

    @property
    def LD(self):
        return self._matrix()

    @LD.setter
    def LD(self, LD):
        self._drop_device_factor()
        if LD is not None:
            self._LD = util.as_matrix(_real_f64_matrix(np.asarray(LD), 'LD'))
        else:
            self.__dict__.pop('_LD', None)

    @property
    def d(self):
        d = self._matrix_diagonal()
        if np.iscomplexobj(d) and np.all(np.isreal(d)):
            d = d.real
        return d

    @property
    def D(self):
        import scipy.sparse
        return scipy.sparse.diags(self.d)

    @property
    def L(self):
        L = self.LD.copy()
        L[np.diag_indices(self.n)] = 1
        return L

    def is_equal(self, other):
        return super().is_equal(other) and util.is_equal(self.LD, other.LD)

    def is_almost_equal(self, other, rtol=1e-04, atol=1e-06):
        return super().is_almost_equal(other) and util.is_almost_equal(self.LD, other.LD, rtol=rtol, atol=atol)

    def as_LDL_Decomposition(self):
        return LDL_Decomposition(self.L, self.d, p=self.p)

    def as_type(self, type_str, copy=False):
        try:
            return super().as_type(type_str, copy=copy)
        except errors.NoDecompositionConversionImplementedError:
            if type_str == constants.LDL_DECOMPOSITION_TYPE:
                return self.as_LDL_Decomposition()
            if type_str == constants.LL_DECOMPOSITION_TYPE:
                return self.as_LDL_Decomposition().as_LL_Decomposition()
            raise

    def is_finite(self):
        return util.is_finite(self.LD)

    def is_positive_semidefinite(self):
        return bool(np.all(self.d >= 0))

    def is_positive_definite(self):
        return bool(np.all(self.d >= np.finfo(self.d.dtype).resolution))

    def is_singular(self):
        return bool(np.any(np.abs(self.d) < np.finfo(self.d.dtype).resolution))

    @property
    def _attribute_names(self):
        return ('LD', 'p')

    def _upload(self, ctx):
        return self.as_LDL_Decomposition()._upload(ctx)  # decompositions.py:1150-1160 convert first, too

    def append_block_decomposition(self, dec):
        dec = self.as_same_type(dec, copy=False)
        return LDL_DecompositionCompressed(p=np.concatenate((self.p, dec.p + self.n)),
                                           LD=util.block_diag(self.LD, dec.LD))


class LL_Decomposition(DecompositionBase):
    """ `L L^H` (Cholesky) of the permuted matrix  (decompositions.py:1170-1373). """

    type_str = constants.LL_DECOMPOSITION_TYPE
    _matrix_attr = '_L'
    _matrix_type_code = _native.TYPE_LL

    def __init__(self, L=None, p=None):
        self.L = L
        super().__init__(p=p)

    @property
    def n(self):
        return self._matrix_n()

    @property
    def L(self):
        return self._matrix()

    @L.setter
    def L(self, L):
        self._drop_device_factor()
        if L is not None:
            self._L = util.as_matrix(_real_f64_matrix(np.asarray(L), 'L'))
        else:
            self.__dict__.pop('_L', None)

    def is_equal(self, other):
        return super().is_equal(other) and util.is_equal(self.L, other.L)

    def is_almost_equal(self, other, rtol=1e-04, atol=1e-06):
        return super().is_almost_equal(other) and util.is_almost_equal(self.L, other.L, rtol=rtol, atol=atol)

    @property
    def _d(self):
        d = self._matrix_diagonal()
        if np.iscomplexobj(d) and np.all(np.isreal(d)):
            d = d.real
        return d

    def as_LDL_Decomposition(self):  # decompositions.py:1248-1284
        L, p, n = self.L, self.p, self.n
        d = self._d
        zero = d == 0
        d_inverse = np.zeros_like(d)
        d_inverse[~zero] = 1 / d[~zero]
        if np.any(zero):
            for i in np.where(zero)[0]:
                if not np.all(np.isclose(L[i + 1:, i], 0)):
                    raise errors.NoDecompositionPossibleWithProblematicSubdecompositionError(
                        self, constants.LDL_DECOMPOSITION_TYPE, p[i])
        L = L * d_inverse[np.newaxis, :]
        L[np.diag_indices(n)] = 1
        return LDL_Decomposition(L, d * d.conj(), p=p)

    def as_type(self, type_str, copy=False):
        try:
            return super().as_type(type_str, copy=copy)
        except errors.NoDecompositionConversionImplementedError:
            if type_str == constants.LDL_DECOMPOSITION_TYPE:
                return self.as_LDL_Decomposition()
            if type_str == constants.LDL_DECOMPOSITION_COMPRESSED_TYPE:
                return self.as_LDL_Decomposition().as_LDL_DecompositionCompressed()
            raise

    def is_finite(self):
        return util.is_finite(self.L)

    def is_positive_semidefinite(self):
        return True

    def is_positive_definite(self):
        return bool(np.all(self._d >= np.finfo(self._d.dtype).resolution))

    def is_singular(self):  # decompositions.py:1312-1314
        return bool(np.any(np.abs(self._d) < np.finfo(self._d.dtype).resolution))

    @property
    def _attribute_names(self):
        return ('L', 'p')

    def _upload(self, ctx):
        L = np.ascontiguousarray(self.L, dtype=np.float64)
        p = np.ascontiguousarray(self.p, dtype=np.int64)
        import ctypes
        h = ctypes.c_void_p()
        rc = ctx.lib.mdb_factor_upload(ctx.handle, self.n, _native.TYPE_LL, _native.ptr(L), self.n, None,
                                       _native.ptr(p), ctypes.byref(h))
        ctx.check(rc, 'mdb_factor_upload')
        return _native.Factor(ctx, h)

    def append_block_decomposition(self, dec):
        dec = self.as_same_type(dec, copy=False)
        return LL_Decomposition(p=np.concatenate((self.p, dec.p + self.n)), L=util.block_diag(self.L, dec.L))


# *** module functions (decompositions.py:1378-1421) *** #

def save(filename, decomposition):
    decomposition.save(filename)


def load(filename):
    type_str = DecompositionBase._load_type(filename)
    for cls in (LDL_Decomposition, LDL_DecompositionCompressed, LL_Decomposition):
        if cls.type_str == type_str:
            decomposition = cls()
            decomposition.load(filename)
            return decomposition
    raise errors.DecompositionInvalidFile(filename)
