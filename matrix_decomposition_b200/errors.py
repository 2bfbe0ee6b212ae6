"""Exception hierarchy of the public API.

The class names, constructor arguments and attributes are those of the reference (matrix/errors.py:8-176), so that
`except matrix.errors.X` clauses and code that reads `.matrix`, `.decomposition`, `.filename`,
`.problematic_leading_principal_submatrix_index` ... keep working.  Here every class only declares the text of its
message; the two bases below fill it in.
"""


class BaseError(Exception):
    """ Root of all exceptions of this package. """

    def __init__(self, message):
        self.message = message
        super().__init__(message)

    def __str__(self):
        return self.message


class _AboutMatrix(BaseError):
    text = 'Error with matrix with shape {shape}'

    def __init__(self, matrix, message=None):
        self.matrix = matrix
        super().__init__(message if message is not None else self.text.format(shape=getattr(matrix, 'shape', None)))


class _AboutDecomposition(BaseError):
    text = 'Error with decomposition {dec}'

    def __init__(self, decomposition, message=None):
        self.decomposition = decomposition
        super().__init__(message if message is not None else self.text.format(dec=decomposition))


# ---- something is wrong with a matrix argument (errors.py:21-76) ----

class MatrixError(_AboutMatrix):
    pass


class MatrixNotSquareError(MatrixError):
    text = 'Matrix with shape {shape} is not a square matrix.'


class MatrixNotFiniteError(MatrixError):
    text = 'Matrix with shape {shape} has not finite entries.'


class MatrixSingularError(MatrixError):
    text = 'Matrix with shape {shape} is singular.'


class MatrixNotHermitianError(MatrixError):
    text = 'Matrix with shape {shape} is not Hermitian.'

    def __init__(self, matrix, i=None, j=None):
        where = ''
        if i is not None and j is not None:
            if i == j:
                where = ' Its {}-th diagonal value {} is complex.'.format(i, matrix[i, i])
            else:
                where = ' Its value at index ({0}, {1}) is {2} and its value at index ({1}, {0}) is {3}.'.format(
                    i, j, matrix[i, j], matrix[j, i])
        super().__init__(matrix, self.text.format(shape=matrix.shape) + where)


class MatrixComplexDiagonalValueError(MatrixNotHermitianError):
    def __init__(self, matrix, i=None):
        super().__init__(matrix, i, i)


# ---- something is wrong with a decomposition object or file (errors.py:81-122) ----

class DecompositionError(_AboutDecomposition):
    pass


class DecompositionNotFiniteError(DecompositionError):
    text = 'Decomposition {dec} has non-finite entries.'


class DecompositionSingularError(DecompositionError):
    text = 'Decomposition {dec} represents a singular matrix.'


class DecompositionInvalidFile(DecompositionError, OSError):
    text = 'File {filename} is not a valid decomposition file.'

    def __init__(self, filename, **details):
        self.filename = filename
        DecompositionError.__init__(self, None, self.text.format(filename=filename, **details))


class DecompositionInvalidDecompositionTypeFile(DecompositionInvalidFile):
    text = 'File {filename} contains a decomposition of type {found} but type {needed} is needed.'

    def __init__(self, filename, type_file, type_needed):
        super().__init__(filename, found=type_file, needed=type_needed)


# ---- a decomposition cannot be computed or converted (errors.py:127-157) ----

class NoDecompositionPossibleError(BaseError):
    def __init__(self, base, desired_type):
        self.base, self.desired_type = base, desired_type
        super().__init__('Base {} can not be converted to type {}.'.format(base, desired_type))


class NoDecompositionPossibleWithProblematicSubdecompositionError(NoDecompositionPossibleError):
    """ Raised by `decompose` for a matrix that is not positive definite: the index of the first leading principal
    submatrix that fails, and (when available) the decomposition computed up to there. """

    def __init__(self, base, desired_type, problematic_leading_principal_submatrix_index, subdecomposition=None):
        super().__init__(base, desired_type)
        self.problematic_leading_principal_submatrix_index = problematic_leading_principal_submatrix_index
        if subdecomposition is not None:
            self.subdecomposition = subdecomposition


class NoDecompositionPossibleTooManyEntriesError(NoDecompositionPossibleError):
    pass


class NoDecompositionConversionImplementedError(NoDecompositionPossibleError):
    pass


# ---- iterative helpers (errors.py:162-176; raised by nothing on the dense path, kept for `except` clauses) ----

class TooManyIterationsError(BaseError):
    """ An iterative calculation gave up.  `iteration` / `result` exist only when they were passed. """

    def __init__(self, message=None, iteration=None, result=None):
        if message is None:
            message = ('No solution found in maximal number of iterations.' if iteration is None
                       else 'No solution found in {} iterations.'.format(iteration))
        super().__init__(message)
        for name, value in (('iteration', iteration), ('result', result)):
            if value is not None:
                setattr(self, name, value)


# ---- this package only ----

class NativeLibraryError(BaseError):
    """ libmdb200.so is missing or a CUDA call failed.  There is no CPU fallback. """
